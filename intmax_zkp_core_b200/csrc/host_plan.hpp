// Host-side planning shared by the C-ABI library and the g++ emulation harness (tests/emu): Goldilocks
// helpers for twiddle-table generation and the pass schedule of a multi-pass transform.  Pure C++, no
// CUDA calls, no kernels; the transforms themselves only exist as CUDA kernels (ntt_kernels.cuh).
// Replaces plonky2_field `fft_root_table` (field/src/fft.rs, plonky2 @ f99ed9c).
#pragma once
#include <cstdint>
#include <vector>

#include "ntt_kernels.cuh"

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

struct TwoLevelPtr {
    const u64* lo;
    const u64* hi;
    u32 lo_bits;
};

struct Plan {
    u32 n_passes;
    u32 bits[MAX_PASSES];
    PassParams pass[MAX_PASSES];
};

// P = ceil(n_log / 8) passes of near-equal width
static inline void split_passes(u32 n_log, u32* bits, u32* n_passes) {
    u32 P = (n_log + 7) / 8;
    *n_passes = P;
    for (u32 i = 0; i < P; i++) bits[i] = n_log / P + (i < n_log % P ? 1 : 0);
}

// wtab[B] = small_root_table(B, dir) on the device; tw = two-level powers of w_n^(+-1) (needed when P > 1)
//   bitrev_out: every pass in place in `out` (DIF order); else natural order through `scratch`
static inline void make_plan(Plan* plan, const u64* in, u64 in_stride, u64* out, u64 out_stride, u64* scratch,
                             u32 n_log, u32 ncols, u64* const* wtab, TwoLevelPtr tw, bool bitrev_out,
                             const TwoLevelPtr* scale, u64 out_scale, bool canon_in) {
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
        p.wtab = wtab[B];
        bool last = (pi + 1 == P);
        if (pi == 0) {
            p.in = in; p.in_col_stride = in_stride;
            p.canon_in = canon_in;
            if (scale) { p.sc_lo = scale->lo; p.sc_hi = scale->hi; p.sc_lo_bits = scale->lo_bits; }
        }
        if (bitrev_out) {
            if (pi > 0) { p.in = out; p.in_col_stride = out_stride; }
            p.out = out; p.out_col_stride = out_stride;
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
        if (!last) { p.tw_lo = tw.lo; p.tw_hi = tw.hi; p.tw_lo_bits = tw.lo_bits; }
        if (last) p.out_scale = out_scale;
        plan->pass[pi] = p;
    }
}

}  // namespace ntt
