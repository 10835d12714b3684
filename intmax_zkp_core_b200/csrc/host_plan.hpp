// Host-side planning shared by the C-ABI library and the g++ emulation harness (tests/emu): Goldilocks
// helpers for twiddle-table generation and the pass schedule of a multi-pass transform.  Pure C++, no
// CUDA calls, no kernels; the transforms themselves only exist as CUDA kernels (ntt_kernels.cuh).
// Replaces plonky2_field `fft_root_table` (field/src/fft.rs, plonky2 @ f99ed9c).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "ntt_kernels.cuh"
#include "ntt_ct_kernels.cuh"

namespace hostgl {
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned __int128 u128;
static const u64 P = 0xFFFFFFFF00000001ull;
static inline u64 mul(u64 a, u64 b) { return (u64)(((u128)a * b) % P); }
static inline u64 pw(u64 b, u64 e) {
    u64 r = 1;
    while (e) { if (e & 1) r = mul(r, b); b = mul(b, b); e >>= 1; }
    return r;
}
static inline u64 inv(u64 a) { return pw(a, P - 2); }
static const u64 G2 = 1753635133440165772ull;  // 7^((p-1)/2^32): generator of the 2^32 subgroup (A10)
static inline u64 root(u32 n_log) { u64 w = G2; for (u32 i = n_log; i < 32; i++) w = mul(w, w); return w; }
static inline u32 bitrev(u32 x, u32 bits) { u32 r = 0; for (u32 i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// w_{2^B}^(+-e) for e < 2^(B-1)
static inline std::vector<u64> small_root_table(u32 B, int dir) {
    u64 w = root(B);
    if (dir) w = inv(w);
    size_t cnt = B ? ((size_t)1 << (B - 1)) : 1;
    std::vector<u64> h(cnt);
    u64 x = 1;
    for (size_t i = 0; i < cnt; i++) { h[i] = x; x = mul(x, w); }
    return h;
}
// base^e for e < 2^bits as lo[e & mask] * hi[e >> lo_bits]; returns lo_bits
static inline u32 two_level_powers(u64 base, u32 bits, std::vector<u64>* lo, std::vector<u64>* hi) {
    u32 lo_bits = (bits + 1) / 2, hi_bits = bits - lo_bits;
    lo->assign((size_t)1 << lo_bits, 0);
    hi->assign((size_t)1 << hi_bits, 0);
    u64 x = 1;
    for (size_t i = 0; i < lo->size(); i++) { (*lo)[i] = x; x = mul(x, base); }
    u64 step = x;  // base^(2^lo_bits)
    x = 1;
    for (size_t i = 0; i < hi->size(); i++) { (*hi)[i] = x; x = mul(x, step); }
    return lo_bits;
}
// leaf block b of the LDE (leaves [b*n, (b+1)*n)) is the coset s_b * <w_n>, s_b = 7 * w_N^bitrev(b, rate_bits)
static inline u64 coset_shift_of_block(u32 n_log, u32 rate_bits, u32 b) {
    return mul(7, pw(root(n_log + rate_bits), bitrev(b, rate_bits)));
}
}  // namespace hostgl

namespace ntt {

struct Plan {
    u32 n_passes;
    u32 bits[MAX_PASSES];
    PassParams pass[MAX_PASSES];
};

// One pass up to 2^10 points (a single launch for small circuits and FRI layers), otherwise P = ceil(n_log / 8) passes of
// near-equal width.  Measured on B200 at 2^20: 10 + 10 (4096-element tiles, 4 batches per tile) loses to 7 + 7 + 6
// (LDE 26.1 vs 23.7 ms): the saved load / store / twiddle product does not pay for the 32-byte access runs.
static inline void split_passes(u32 n_log, u32* bits, u32* n_passes) {
    u32 P = n_log <= (u32)MAX_PASS_BITS ? (n_log ? 1 : 0) : (n_log + 7) / 8;
    *n_passes = P;
    for (u32 i = 0; i < P; i++) bits[i] = n_log / P + (i < n_log % P ? 1 : 0);
}

// What a transform needs from the caller's table cache.
struct TransformTables {
    u64* const* wtab;          // [B] -> small_root_table(B, dir) on the device
    const u64* twimg[MAX_PASSES];   // per non-final pass: twiddle image (nullptr for the final pass)
    const u64* scale;          // optional [n_blk][n] input scaling tables (coset shift powers)
    u64 scale_blk_stride;
};

// geometry of the twiddle image of pass `pi`: entries = 2^(B + C_log), exponent shift = n_log - B - C_log
struct TwiddleImageShape { u32 B, C_log, shift; };
static inline u32 twiddle_images(u32 n_log, TwiddleImageShape* shapes) {
    u32 bits[MAX_PASSES], P;
    split_passes(n_log, bits, &P);
    u32 consumed = 0;
    for (u32 pi = 0; pi + 1 < P; pi++) {
        consumed += bits[pi];
        shapes[pi] = TwiddleImageShape{bits[pi], n_log - consumed, consumed - bits[pi]};
    }
    return P ? P - 1 : 0;
}

//   bitrev_out: every pass in place in `out` (DIF order); else natural order through `scratch`
//   out_scale (n^-1 of an inverse transform) must already be folded into twimg[0] when P > 1
static inline void make_plan(Plan* plan, const u64* in, u64 in_stride, u64* out, u64 out_stride, u64* scratch,
                             u32 n_log, u32 ncols, const TransformTables& tb, bool bitrev_out, u64 out_scale,
                             bool canon_in, u64 in_blk_stride, u64 out_blk_stride) {
    u32 P;
    split_passes(n_log, plan->bits, &P);
    plan->n_passes = P;
    u32 consumed = 0;
    for (u32 pi = 0; pi < P; pi++) {
        u32 B = plan->bits[pi];
        consumed += B;
        PassParams p{};
        p.ncols = ncols;
        p.n_log = n_log;
        p.C_log = n_log - consumed;
        p.wtab = tb.wtab[B];
        bool last = (pi + 1 == P);
        if (pi == 0) {
            p.in = in; p.in_col_stride = in_stride; p.in_blk_stride = in_blk_stride;
            p.canon_in = canon_in;
            p.scale = tb.scale; p.scale_blk_stride = tb.scale_blk_stride;
        }
        if (bitrev_out) {
            if (pi > 0) { p.in = out; p.in_col_stride = out_stride; p.in_blk_stride = out_blk_stride; }
            p.out = out; p.out_col_stride = out_stride; p.out_blk_stride = out_blk_stride;
            p.out_mode = OUT_INPLACE_BITREV;
        } else {
            if (pi > 0) { p.in = scratch; p.in_col_stride = (u64)1 << n_log; }
            if (last) {
                p.out = out; p.out_col_stride = out_stride;
                p.out_mode = OUT_FINAL_NATURAL;
                p.n_digits = P - 1;
                for (u32 q = 0; q + 1 < P; q++) p.digits[q] = plan->bits[q];
            } else {
                p.out = scratch; p.out_col_stride = (u64)1 << n_log;
                p.out_mode = OUT_INPLACE_NATURAL;
            }
        }
        if (!last) p.twimg = tb.twimg[pi];
        if (last && P == 1) p.out_scale = out_scale;
        plan->pass[pi] = p;
    }
}

}  // namespace ntt

// ---- second-generation passes (ntt_ct_kernels.cuh): schedule of one multi-pass transform
namespace ntc {

struct Plan {
    u32 n_passes;
    u32 bits[ntt::MAX_PASSES];
    int kind[ntt::MAX_PASSES];
    PassParams pass[ntt::MAX_PASSES];
    u64 grid[ntt::MAX_PASSES];      // CTAs of the launch
};

// does this generation cover a transform of 2^n_log points?  (two or more passes of 5..8 bits, strided tiles at least T wide)
static inline bool covers(u32 n_log) {
    if (n_log <= (u32)ntt::MAX_PASS_BITS || n_log > 32) return false;
    u32 bits[ntt::MAX_PASSES], P, S = 0;
    ntt::split_passes(n_log, bits, &P);
    if (P < 2 || P > (u32)ntt::MAX_PASSES) return false;
    for (u32 pi = 0; pi < P; pi++) {
        if (bits[pi] < (u32)MIN_BITS || bits[pi] > (u32)MAX_BITS) return false;
        S += bits[pi];
        if (pi + 1 < P && n_log - S < 11 - bits[pi]) return false;      // C >= T
    }
    return true;
}

// Z[i] of the transform on the coset s <w>, i in [1, n): host replica of build_ztab_kernel (emulation harness, tests)
static inline std::vector<u64> ztab_host(u32 n_log, u64 s, int dir, u64 last_scale) {
    u64 n = (u64)1 << n_log;
    u64 w = hostgl::root(n_log);
    if (dir) w = hostgl::inv(w);
    std::vector<u64> spow(n_log), z(n, 0);
    u64 x = s % hostgl::P;
    for (u32 e = 0; e < n_log; e++) { spow[e] = x; x = hostgl::mul(x, x); }
    for (u64 i = 1; i < n; i++) {
        u32 l = 63 - (u32)__builtin_clzll(i);
        u32 j = (u32)(i - ((u64)1 << l));
        u64 e = (u64)hostgl::bitrev(j, l) << (n_log - 1 - l);
        u64 r = hostgl::mul(hostgl::pw(w, e), spow[n_log - 1 - l]);
        if (last_scale && l == n_log - 1) r = hostgl::mul(r, last_scale);
        z[i] = r;
    }
    return z;
}

//   natural_out: first pass in -> scratch, middle passes in place in scratch, last pass scatters scratch -> out (inverse
//   transform; a_scale = n^-1 and the Z table's last level carries the same factor); otherwise in -> out, then in place in
//   out, bit-reversed (leaf) order.  n_blk > 1: every block reads the same input (in_blk_stride = 0 in the first pass) and
//   writes at out + blk * out_blk_stride.
// width of the last pass / entries per block of the strided table / words per block of the tile-ordered last-pass table
static inline u32 last_pass_bits(u32 n_log) {
    u32 bits[ntt::MAX_PASSES], P;
    ntt::split_passes(n_log, bits, &P);
    return bits[P - 1];
}
static inline u32 first_pass_bits(u32 n_log) {
    u32 bits[ntt::MAX_PASSES], P;
    ntt::split_passes(n_log, bits, &P);
    return bits[0];
}
static inline u64 ztab_entries(u32 n_log) { return (u64)1 << (n_log - last_pass_bits(n_log)); }
static inline u64 zfinal_words(u32 n_log) {
    const u32 B = last_pass_bits(n_log);
    return (((u64)1 << n_log) >> 11) * ((u64)(TILE >> B) * (((u64)1 << B) + 1));
}
// host replica of build_zfinal_kernel from the full table of ztab_host
static inline std::vector<u64> zfinal_host(u32 n_log, const std::vector<u64>& z, bool natural) {
    const u32 B = last_pass_bits(n_log), T_log = 11 - B, TWP = (1u << B) + 1, S = n_log - B;
    std::vector<u64> f(zfinal_words(n_log), 0);
    for (u64 x = 0; x < f.size(); x++) {
        const u64 tile_i = x / ((u64)TWP << T_log);
        const u32 rem = (u32)(x - tile_i * ((u64)TWP << T_log));
        const u32 b = rem / TWP, e = rem - b * TWP;
        if (e < 1 || e >= (1u << B)) continue;
        const u32 batch = (u32)(tile_i << T_log) + b;
        const u32 a_b = natural ? hostgl::bitrev(batch, S) : batch;
        const u32 u = 31 - (u32)__builtin_clz(e);
        f[x] = z[(((((u64)1 << S) + a_b)) << u) + (e ^ (1u << u))];      // ntc::z_index
    }
    return f;
}

// The columns of one launch sequence of a partitioned LDE (sharded.inl); see PassParams.  `count` grid columns: a range of
// physical columns from col0 (sel = 0) or the first `count` columns of source `src_rank` from group col0 / period on (sel = 1).
// pull: the first pass reads every column from its owner's matrix src[q] (local columns src_col_stride apart) and also stores
// it to `copy_out`.
struct ColumnSet {
    u32 run = 0, period = 0, col0 = 0, limit = 0, count = 0, sel = 0, src_rank = 0;
    bool pull = false;
    const u64* src[MAX_SRC] = {};
    u64 src_col_stride = 0;
    u64* copy_out = nullptr;
    u64 copy_col_stride = 0;
};

// ztab / zfinal: the strided table (ztab_entries per block) and the tile-ordered last-pass table (zfinal_words per block)
// cols (optional): see ColumnSet; `ncols` is then ignored (cols->count grid columns) and `in` is only read when !cols->pull
static inline bool make_plan(Plan* plan, const u64* in, u64 in_stride, u64* out, u64 out_stride, u64* scratch, u32 n_log,
                             u32 ncols, u32 n_blk, u64 out_blk_stride, bool natural_out, const u64* ztab, const u64* zfinal,
                             u64 a_scale, bool use_tma, const ColumnSet* cols = nullptr) {
    if (cols) {
        if (natural_out || cols->run == 0 || cols->period == 0 || cols->period % cols->run || cols->count == 0) return false;
        if (cols->pull && (n_blk > (u32)MAX_LOOP_BLOCKS || cols->period / cols->run > (u32)MAX_SRC || cols->sel)) return false;
        if (!cols->sel && (u64)cols->col0 + cols->count > cols->limit) return false;      // (the persistent gather has no idle CTAs)
        ncols = cols->count;
    }
    if (!covers(n_log) || ncols == 0 || n_blk == 0) return false;
    if (natural_out && (!scratch || n_blk != 1)) return false;
    u32 P;
    ntt::split_passes(n_log, plan->bits, &P);
    plan->n_passes = P;
    const u64 n = (u64)1 << n_log;
    u64* work = natural_out ? scratch : out;
    const u64 work_stride = natural_out ? n : out_stride;
    u32 S = 0;
    for (u32 pi = 0; pi < P; pi++) {
        const u32 B = plan->bits[pi];
        const bool last = pi + 1 == P;
        PassParams p{};
        p.ncols = ncols; p.n_blk = n_blk; p.n_log = n_log; p.S = S; p.C_log = n_log - S - B;
        p.ztab = last ? zfinal : ztab;
        p.ztab_blk_stride = last ? zfinal_words(n_log) : ztab_entries(n_log);
        p.use_tma = use_tma ? 1u : 0u;
        if (pi == 0) { p.in = in; p.in_col_stride = in_stride; p.in_blk_stride = 0; }
        else { p.in = work; p.in_col_stride = work_stride; p.in_blk_stride = out_blk_stride; }
        if (last && natural_out) { p.out = out; p.out_col_stride = out_stride; p.out_blk_stride = 0; p.a_scale = a_scale; }
        else { p.out = work; p.out_col_stride = work_stride; p.out_blk_stride = out_blk_stride; }
        int kind;
        if (last) kind = natural_out ? KIND_FINAL_NATURAL : KIND_FINAL_INPLACE;
        else kind = (pi == 0 && n_blk > 1 && n_blk <= (u32)MAX_LOOP_BLOCKS) ? KIND_STRIDED_LOOP : KIND_STRIDED;
        if (cols) {
            p.col_run = cols->run; p.col_period = cols->period; p.col0 = cols->col0; p.col_limit = cols->limit;
            p.col_sel = cols->sel; p.col_src = cols->src_rank;
            if (pi == 0 && cols->pull) {
                kind = KIND_PULL_LOOP;
                p.in = nullptr; p.in_col_stride = cols->src_col_stride;
                p.copy_out = cols->copy_out; p.copy_col_stride = cols->copy_col_stride;
                for (u32 q = 0; q < cols->period / cols->run; q++) {
                    p.src[q] = cols->src[q];
                    if ((uintptr_t)cols->src[q] & 15) p.use_tma = 0;
                }
                if (((uintptr_t)p.copy_out & 15) || (p.copy_col_stride & 1)) p.use_tma = 0;
            }
        }
        // TMA needs 16-byte aligned bases and strides, and a real (non-zero) block stride wherever a block coordinate moves
        if (kind == KIND_STRIDED || kind == KIND_STRIDED_LOOP || kind == KIND_PULL_LOOP) {
            const bool in_blk_moves = kind == KIND_STRIDED && n_blk > 1;
            if (((uintptr_t)p.in | (uintptr_t)p.out) & 15) p.use_tma = 0;
            if ((p.in_col_stride | p.out_col_stride | p.out_blk_stride) & 1) p.use_tma = 0;
            if (in_blk_moves && (p.in_blk_stride == 0 || (p.in_blk_stride & 1))) p.use_tma = 0;
            if (n_blk > 1 && p.out_blk_stride == 0) p.use_tma = 0;
        } else {
            p.use_tma = 0;
            if ((((uintptr_t)p.in | (uintptr_t)p.out) & 15) || ((p.in_col_stride | p.out_col_stride | p.in_blk_stride | p.out_blk_stride) & 1))
                return false;       // the last pass moves pairs of elements (128-bit accesses)
        }
        plan->kind[pi] = kind;
        plan->pass[pi] = p;
        plan->grid[pi] = (n >> 11) * (u64)ncols * ((kind == KIND_STRIDED_LOOP || kind == KIND_PULL_LOOP) ? 1 : n_blk);
        S += B;
    }
    return true;
}

}  // namespace ntc
