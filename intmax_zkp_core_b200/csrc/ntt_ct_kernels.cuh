// Goldilocks NTT, second generation: block-twiddle Cooley-Tukey passes for transforms of 2^11 points and more.
//
// Replaces the same plonky2_field functions as ntt_kernels.cuh (`fft_classic`, `ifft_with_options`,
// `coset_fft_with_options`, `lde`, plus plonky2 `transpose` + `reverse_index_bits_in_place`; plonky2 @ f99ed9c,
// field/src/fft.rs, field/src/polynomial/mod.rs, plonky2/src/fri/oracle.rs `lde_values`),
// reached from the reference through PolynomialBatch::from_values / from_coeffs in every prove() / build(), e.g.
// /root/reference/src/rollup/circuits/mod.rs:605,1247.  ntt_kernels.cuh keeps the small sizes (one pass, n <= 2^10) and
// the callers that bring their own scale tables (FRI layers).
//
// The network.  A transform of n = 2^L coefficients on the coset s<w_n>, natural order in, bit-reversed order out:
//   level l = 0..L-1 cuts the array into 2^l blocks of m = n / 2^l positions; block j evaluates on the coset
//   sigma = s * w_n^bitrev_l(j) and its butterflies are   t = z * x[i + m/2];  x[i] += t;  x[i + m/2] = x[i] - t
//   with ONE twiddle per block,  z = sigma^(m/2) = s^(n / 2^(l+1)) * w_(2^(l+1))^bitrev_l(j)   (table Z[2^l + j]).
// Against the four-step form of ntt_kernels.cuh this removes, per element, the coset-shift product (the shift lives in Z)
// and the twiddle product between passes (there is none), and a radix-8 group reads 7 twiddles instead of 12.
// The passes are bound by the integer pipes, not by HBM (10 butterflies of ~28 instructions per 8-byte element), so the
// instruction count is what is optimised:
//   * lazy arithmetic: values travel as arbitrary u64; only the product is canonicalised (a carry chain, gl::canon_cc),
//     sums and differences fold one wrap (gl::add_nc / gl::sub_nc); the last pass canonicalises what it stores;
//   * the first pass of an LDE stages its coefficient tile ONCE and produces all 2^rate_bits cosets from it;
//   * strided tiles ([2^B rows][T contiguous elements], row pitch C) move with TMA: one cp.async.bulk.tensor per tile and
//     direction through a 5-D tensor map (c, row, a, coset block, column), completion on an mbarrier / bulk group, so no
//     address arithmetic, LDG or STS of the staging is left in the issue stream; the dense [row][T] image has its T
//     elements on consecutive lanes, which is conflict free without padding.
//   * the inverse transform scatters its last pass to natural order in runs of T elements (bit reversal of the batch
//     index, as plonky2's reverse_index_bits), with n^-1 folded into the last level.
// Every pass is in place except the first (which may read another buffer) and the natural-order scatter.
#pragma once
#include "goldilocks.cuh"

#ifndef B200ZKP_HOST_EMU
#include <cuda.h>   // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint by the host code)
#endif

namespace ntc {

using gl::u32;
using gl::u64;

static constexpr int TILE = 2048;      // elements per tile
static constexpr int THREADS = 256;    // 8 elements per thread
static constexpr int MIN_BITS = 5, MAX_BITS = 8;      // pass widths this generation is instantiated for
static constexpr int MAX_LOOP_BLOCKS = 8;             // coset blocks one CTA produces from one staged tile
static constexpr int MAX_SRC = 8;                     // sources (ranks of one NVLink domain) of a pull pass

enum Kind : int {
    KIND_STRIDED = 0,        // [A][2^B][C] view, C >= T: tile = 2^B rows x T contiguous elements; coset blocks are grid columns
    KIND_STRIDED_LOOP = 1,   // same tile, read once, every coset block produced by the same CTA (first pass of an LDE)
    KIND_FINAL_INPLACE = 2,  // C = 1: tile = T batches of 2^B contiguous elements, stored where they came from (LDE leaf order)
    KIND_FINAL_NATURAL = 3,  // C = 1: batches picked by bit-reversed index, stored in natural order (inverse transform)
    KIND_PULL_LOOP = 4,      // KIND_STRIDED_LOOP whose coefficient tiles come from per-source matrices (peer memory) and are
                             // copied to a local matrix on the way: all-gather + first pass of a partitioned LDE in one kernel
};

struct PassParams {
    const u64* in;
    u64* out;
    u64 in_col_stride, out_col_stride;    // elements between columns
    u64 in_blk_stride, out_blk_stride;    // elements between the coset blocks of a column (0: every block reads the same input)
    u32 ncols, n_blk;
    u32 n_log;      // L
    u32 S;          // levels done by earlier passes
    u32 C_log;      // L - S - B
    const u64* ztab;          // strided passes: Z of block b at ztab + b * ztab_blk_stride, entries [1, 2^(L - B_last))
    u64 ztab_blk_stride;      // final passes: the tile-ordered image of the last levels (see FinalSmem), block stride in words
    u64 a_scale;    // 0, or n^-1: multiplies the sum operand of the LAST level (its twiddles carry the same factor)
    u32 use_tma;
    // column sets of a partitioned LDE (sharded.inl).  col_run = 0: grid column v is column v of in / out.  Otherwise the
    // columns of the commitment are dealt to the ranks in groups: column c belongs to source q = (c % col_period) / col_run
    // and is that source's local column (c / col_period) * col_run + c % col_run (col_period = col_run * number of ranks, a
    // multiple of the sponge rate, so that every group of col_period columns can be hashed as soon as it is complete).
    //   col_sel = 0: grid column v is physical column col0 + v (a range of the commitment's columns, all sources)
    //   col_sel = 1: grid column v is the v-th column of source col_src only: physical col0 + (v / col_run) * col_period +
    //                col_src * col_run + v % col_run  (coset transforms of one shard that has arrived)
    // A CTA whose physical column is >= col_limit leaves (ragged last group).
    u32 col_run, col_period, col0, col_limit, col_sel, col_src;
    // pull pass (KIND_PULL_LOOP): source q's coefficients are its local columns of the matrix at src[q] (in_col_stride apart;
    // a peer GPU's exchange window, read over NVLink), and the staged tile is also stored to the physical column of copy_out,
    // so that the all-gather of the coefficients happens tile by tile inside the transform
    const u64* src[MAX_SRC];
    u64* copy_out;
    u64 copy_col_stride;
};

// physical column of grid column v, or 0xFFFFFFFF when the CTA has nothing to do; q, i: owning source and its local column
GL_FN u32 map_column(const PassParams& p, u32 v, u32& q, u32& i) {
    if (p.col_run == 0) { q = 0; i = v; return v; }
    u32 phys;
    if (p.col_sel) phys = p.col0 + (v / p.col_run) * p.col_period + p.col_src * p.col_run + v % p.col_run;
    else phys = p.col0 + v;
    const u32 in_grp = phys % p.col_period;
    q = in_grp / p.col_run;
    i = (phys / p.col_period) * p.col_run + in_grp % p.col_run;
    return phys < p.col_limit ? phys : 0xFFFFFFFFu;
}

GL_FN u32 ulog2(u32 e) {
#ifdef B200ZKP_HOST_EMU
    u32 r = 0; while (e >>= 1) r++; return r;
#else
    return 31u - (u32)__clz((int)e);
#endif
}
GL_FN u32 bitrev_bits(u32 x, u32 bits) { return bits ? (gl::brev32(x) >> (32 - bits)) : 0; }

// index into Z of sub-twiddle e (1 <= e < 2^B: level u = floor(log2 e) of the pass, block h = e - 2^u inside the tile) for a
// tile whose rows share the block prefix `prefix` = 2^S + a
GL_FN u64 z_index(u64 prefix, u32 e) {
    const u32 u = ulog2(e);
    return (prefix << u) + (e ^ (1u << u));
}

// t = z * b canonical; (a, b) <- (a + t, a - t), arbitrary u64 representatives
GL_FN void butterfly(u64& a, u64& b, u64 z) {
    const u64 t = gl::canon_cc(gl::mul_nc(b, z));
    const u64 a0 = a;
    a = gl::add_nc(a0, t);
    b = gl::sub_nc(a0, t);
}

// levels u0 .. u0+A-1 of the pass on the 2^A values x[j] <-> row gbase + j * stride; gh = row >> (B - u0) (the rows of one
// group share it), tw = the tile's (or the batch's) sub-twiddle table.  SCALE: the last level multiplies the sum operand.
template <int A, bool SCALE>
GL_FN void ct_group(u64 (&x)[1 << A], const u64* __restrict__ tw, u32 u0, u32 gh, u64 a_scale) {
#pragma unroll
    for (int v = 0; v < A; v++) {
        const int half = (1 << A) >> (v + 1);
        u64 z[1 << (A - 1)];                     // 2^v twiddles at this level
#pragma unroll
        for (int q = 0; q < (1 << v); q++) z[q] = tw[(1u << (u0 + v)) + (gh << v) + q];
#pragma unroll
        for (int j = 0; j < (1 << A); j++) {
            if (j & half) continue;
            if (SCALE && v == A - 1) x[j] = gl::canon_cc(gl::mul_nc(x[j], a_scale));
            butterfly(x[j], x[j + half], z[j >> (A - v)]);
        }
    }
}

// the thread-level pass bodies are written once and compiled two ways (see ntt_kernels.cuh): on the device one CUDA
// thread runs each NTC_FOR_THREADS body; under B200ZKP_HOST_EMU (tests only) a loop steps the 256 threads of the CTA
#ifdef B200ZKP_HOST_EMU
#define NTC_FOR_THREADS(tid) for (u32 tid = 0; tid < (u32)THREADS; tid++)
#define NTC_SYNC() do {} while (0)
#else
#define NTC_FOR_THREADS(tid) for (u32 tid = threadIdx.x, once__ = 1; once__; once__ = 0)
#define NTC_SYNC() __syncthreads()
#endif

// one round: every thread runs 8 >> A groups of 2^A rows.  src/dst: [row * pitch + t] images (src != dst only in the first
// round of the coset loop, which reads the staged coefficients and writes the work tile).  tw_pitch: 0 (one table for the
// tile) or the per-batch table pitch of a final pass (the "column" t is then a batch with its own twiddles).
template <int A, bool SCALE>
GL_FN void run_round(const u64* __restrict__ src, u64* __restrict__ dst, const u64* __restrict__ tw, u32 tw_pitch, u32 B,
                     u32 T_log, u32 pitch, u32 u0, u64 a_scale, u32 tid) {
    const u32 st_log = B - u0 - A;
    constexpr int UNITS = 8 >> A;
#pragma unroll
    for (int r = 0; r < UNITS; r++) {
        const u32 U = tid + r * THREADS;
        const u32 t = U & ((1u << T_log) - 1), w = U >> T_log;
        const u32 gl_ = w & ((1u << st_log) - 1), gh = w >> st_log;
        const u32 gbase = (gh << (st_log + A)) + gl_;
        u64 x[1 << A];
#pragma unroll
        for (int j = 0; j < (1 << A); j++) x[j] = src[(gbase + ((u32)j << st_log)) * pitch + t];
        ct_group<A, SCALE>(x, tw + t * tw_pitch, u0, gh, a_scale);
#pragma unroll
        for (int j = 0; j < (1 << A); j++) dst[(gbase + ((u32)j << st_log)) * pitch + t] = x[j];
    }
}

// all rounds of a B-bit pass: the short round (B mod 3 levels) first, radix-8 rounds after it.  The caller places the barrier
// after the last round (the TMA store path needs its proxy fence between the last writes and that barrier).
template <int B, bool SCALE>
GL_FN void run_rounds(const u64* first_src, u64* tile, const u64* tw, u32 tw_pitch, u32 T_log, u32 pitch, u64 a_scale) {
    u32 u0 = 0;
    const u64* src = first_src;
    if (B % 3 == 2) { NTC_FOR_THREADS(tid) { run_round<2, false>(src, tile, tw, tw_pitch, B, T_log, pitch, u0, a_scale, tid); } u0 += 2; src = tile; NTC_SYNC(); }
    if (B % 3 == 1) { NTC_FOR_THREADS(tid) { run_round<1, false>(src, tile, tw, tw_pitch, B, T_log, pitch, u0, a_scale, tid); } u0 += 1; src = tile; NTC_SYNC(); }
#pragma unroll 1
    for (int r = 0; r < B / 3 - 1; r++) {
        NTC_FOR_THREADS(tid) { run_round<3, false>(src, tile, tw, tw_pitch, B, T_log, pitch, u0, a_scale, tid); }
        u0 += 3; src = tile;
        NTC_SYNC();
    }
    NTC_FOR_THREADS(tid) { run_round<3, SCALE>(src, tile, tw, tw_pitch, B, T_log, pitch, u0, a_scale, tid); }
}

// ---------------------------------------------------------------------------------------------------- TMA / mbarrier
#ifndef B200ZKP_HOST_EMU
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "NTC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NTC_DONE;\n\t"
        "bra NTC_WAIT;\n\t"
        "NTC_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// tile [2^B rows][T] at tensor coordinates (c0, 0, a, blk, col)
__device__ __forceinline__ void tma_load_tile(void* dst, const CUtensorMap* map, u32 c0, u32 a, u32 blk, u32 col, u64* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(0u), "r"(a), "r"(blk), "r"(col), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_tile(const CUtensorMap* map, u32 c0, u32 a, u32 blk, u32 col, const void* src) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];"
                 ::"l"(map), "r"(c0), "r"(0u), "r"(a), "r"(blk), "r"(col), "r"(smem_u32(src)) : "memory");
}
// contiguous run global -> shared (UBLKCP), completion on the mbarrier; 16-byte aligned addresses and size
__device__ __forceinline__ void bulk_load(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// ---------------------------------------------------------------------------------------------------- strided passes
// shared memory image of the strided kernels (dynamic, 128-byte aligned by the launcher's declaration):
//   LOOP: [staged coefficients TILE][work 0 TILE][work 1 TILE][sw: n_blk x 2^B][mbarrier]      others: [tile TILE][sw: 2^B][mbarrier]
template <int B, bool LOOP>
struct StridedSmem {
    static constexpr u32 NPTS = 1u << B;
    static constexpr u32 tiles = LOOP ? 3 : 1;
    static constexpr u32 sw_words = (LOOP ? (u32)MAX_LOOP_BLOCKS : 1u) * NPTS;
    static constexpr u32 bytes = tiles * TILE * 8 + sw_words * 8 + 16;
};

// PULL (LOOP only): tm_in points to MAX_SRC maps (one per source), tm_copy maps the local copy of the coefficients
template <int B, bool LOOP, typename TMap, bool PULL = false>
GL_FN void strided_body(const PassParams& p, const TMap* tm_in, const TMap* tm_out, u64* smem, u32 bid, const TMap* tm_copy = nullptr) {
    static_assert(!PULL || LOOP, "a pull pass stages its tile apart from the work tiles");
    constexpr u32 NPTS = 1u << B;
    constexpr u32 T_log = 11 - B, T = 1u << T_log;
    constexpr u32 IT = TILE / THREADS;
    u64* const stage = smem;                                   // LOOP: the coefficients, staged once; else: the tile
    u64* const work0 = LOOP ? smem + TILE : smem;
    u64* const sw = smem + StridedSmem<B, LOOP>::tiles * TILE;
    u64* const bar = sw + StridedSmem<B, LOOP>::sw_words;

    const u32 V = LOOP ? p.ncols : p.ncols * p.n_blk;
    const u32 vcol = bid % V, tile_i = bid / V;
    u32 src_q, src_i;
    const u32 col = map_column(p, LOOP ? vcol : vcol / p.n_blk, src_q, src_i);
    if (col == 0xFFFFFFFFu) return;                            // (CTA-uniform) a short last source
    const u32 in_col = PULL ? src_i : col;
    const u32 blk_fixed = LOOP ? 0u : vcol % p.n_blk;
    const u32 tiles_per_a_log = p.C_log - T_log;
    const u32 a = tile_i >> tiles_per_a_log;
    const u32 c0 = (tile_i & ((1u << tiles_per_a_log) - 1)) << T_log;
    const u64 prefix = ((u64)1 << p.S) + a;
    const u64 elem0 = ((u64)a << (B + p.C_log)) + c0;            // element (row 0, t = 0) inside a column block
    const u64* __restrict__ in = (PULL ? p.src[src_q] : p.in) + (u64)in_col * p.in_col_stride + (u64)blk_fixed * p.in_blk_stride + elem0;
    (void)tm_in; (void)tm_out; (void)tm_copy; (void)bar;

    // ---- stage the input tile and the sub-twiddles
#ifndef B200ZKP_HOST_EMU
    const bool tma = p.use_tma != 0;
    if (tma) {
        if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, TILE * 8);
            tma_load_tile(stage, PULL ? tm_in + src_q : tm_in, c0, a, blk_fixed, in_col, bar);
        }
    }
#else
    const bool tma = false;
#endif
    NTC_FOR_THREADS(tid) {
        const u32 nb = LOOP ? p.n_blk : 1u;
        for (u32 i = tid; i < nb * NPTS; i += THREADS) {
            const u32 b = LOOP ? i >> B : blk_fixed, e = i & (NPTS - 1);
            if (e) sw[i] = gl::ldg(p.ztab + (u64)b * p.ztab_blk_stride + z_index(prefix, e));
        }
        if (!tma) {
            u64 v[IT];
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                const u32 i = tid + it * THREADS;
                v[it] = in[((u64)(i >> T_log) << p.C_log) + (i & (T - 1))];
            }
#pragma unroll
            for (u32 it = 0; it < IT; it++) stage[tid + it * THREADS] = v[it];
            if (PULL) {
                u64* __restrict__ cp = p.copy_out + (u64)col * p.copy_col_stride + elem0;
#pragma unroll
                for (u32 it = 0; it < IT; it++) {
                    const u32 i = tid + it * THREADS;
                    cp[((u64)(i >> T_log) << p.C_log) + (i & (T - 1))] = v[it];
                }
            }
        }
    }
#ifndef B200ZKP_HOST_EMU
    if (tma) {
        mbar_wait(bar, 0);
        // the gathered tile goes to the local coefficient matrix straight from shared memory (the async proxy wrote it, the
        // async proxy reads it: no fence); the group is the oldest of this CTA and completes under the first coset's rounds
        if (PULL && threadIdx.x == 0) { tma_store_tile(tm_copy, c0, a, 0, col, stage); tma_commit(); }
    }
#endif
    NTC_SYNC();

    // ---- rounds + store, once per coset block
    const u32 n_loop = LOOP ? p.n_blk : 1u;
    for (u32 lb = 0; lb < n_loop; lb++) {
        const u32 blk = LOOP ? lb : blk_fixed;
        u64* const tile = LOOP ? work0 + (lb & 1) * TILE : work0;
#ifndef B200ZKP_HOST_EMU
        // the bulk store issued two blocks ago may still be reading this work tile
        if (LOOP && tma && lb >= 2) { if (threadIdx.x == 0) tma_wait_read<1>(); __syncthreads(); }
#endif
        run_rounds<B, false>(LOOP ? stage : tile, tile, sw + (LOOP ? lb * NPTS : 0), 0, T_log, T, 0);
        u64* __restrict__ out = p.out + (u64)col * p.out_col_stride + (u64)blk * p.out_blk_stride + elem0;
#ifndef B200ZKP_HOST_EMU
        if (tma) {
            fence_proxy_async();                    // the rounds wrote the tile through the generic proxy
            __syncthreads();
            if (threadIdx.x == 0) { tma_store_tile(tm_out, c0, a, blk, col, tile); tma_commit(); }
            continue;
        }
#endif
        NTC_SYNC();
        NTC_FOR_THREADS(tid) {
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                const u32 i = tid + it * THREADS;
                out[((u64)(i >> T_log) << p.C_log) + (i & (T - 1))] = tile[i];
            }
        }
        NTC_SYNC();
    }
#ifndef B200ZKP_HOST_EMU
    if (tma && threadIdx.x == 0) tma_wait_read<0>();       // shared memory must outlive the last bulk store's reads
#endif
}

// ---------------------------------------------------------------------------------------------------- final passes
// shared memory: [tile: 2^B rows x (T + 1)][per-batch sub-twiddles: T x (2^B + 1)][mbarrier]
// The sub-twiddles of a tile are ONE contiguous run of the tile-ordered table (built once per transform shape by
// build_zfinal_kernel): tile i of a block at words [i * tw_words, (i + 1) * tw_words), batch b at b * TWP, sub-twiddle e
// (level u = floor(log2 e), block h = e - 2^u of the batch) at + e; words 0 and 2^B of a batch are padding, so that the
// batches of the 32 lanes of a warp start in different banks.  One cp.async.bulk (UBLKCP) stages it.
template <int B>
struct FinalSmem {
    static constexpr u32 NPTS = 1u << B, T = (u32)TILE >> B;
    static constexpr u32 TP = T + 1, TWP = NPTS + 1;
    static constexpr u32 tw_words = T * TWP;
    static constexpr u32 bytes = (NPTS * TP + tw_words) * 8 + 16;
};

template <int B, bool NATURAL>
GL_FN void final_body(const PassParams& p, u64* smem, u32 bid) {
    constexpr u32 NPTS = 1u << B;
    constexpr u32 T_log = 11 - B, T = 1u << T_log, TP = FinalSmem<B>::TP, TWP = FinalSmem<B>::TWP;
    constexpr u32 IT = TILE / THREADS;
    u64* const tile = smem;
    u64* const swb = smem + NPTS * TP;
    u64* const bar = swb + FinalSmem<B>::tw_words;
    (void)bar;
    const u32 batches_log = p.n_log - B;                       // = S
    const u32 V = p.ncols * p.n_blk;
    const u32 vcol = bid % V, tile_i = bid / V;
    u32 src_q, src_i;
    const u32 col = map_column(p, vcol / p.n_blk, src_q, src_i), blk = vcol % p.n_blk;
    if (col == 0xFFFFFFFFu) return;
    const u32 batch0 = tile_i << T_log;
    const u64* __restrict__ in = p.in + (u64)col * p.in_col_stride + (u64)blk * p.in_blk_stride;
    u64* __restrict__ out = p.out + (u64)col * p.out_col_stride + (u64)blk * p.out_blk_stride;
    const u64* __restrict__ ztile = p.ztab + (u64)blk * p.ztab_blk_stride + (u64)tile_i * FinalSmem<B>::tw_words;

#ifndef B200ZKP_HOST_EMU
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, FinalSmem<B>::tw_words * 8);
        bulk_load(swb, ztile, FinalSmem<B>::tw_words * 8, bar);
    }
#else
    for (u32 i = 0; i < FinalSmem<B>::tw_words; i++) swb[i] = ztile[i];
#endif
    NTC_FOR_THREADS(tid) {
        // data: batch b = rows [a_b << B, (a_b + 1) << B) of the column block; pairs of elements per access
        u64 v[IT];
#pragma unroll
        for (u32 it = 0; it < IT; it += 2) {
            const u32 i = 2 * (tid + (it / 2) * THREADS);
            const u32 b = i >> B, g = i & (NPTS - 1);
            const u32 a_b = NATURAL ? bitrev_bits(batch0 + b, batches_log) : batch0 + b;
#ifdef B200ZKP_HOST_EMU
            v[it] = in[((u64)a_b << B) + g]; v[it + 1] = in[((u64)a_b << B) + g + 1];
#else
            const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(in + ((u64)a_b << B) + g);
            v[it] = q.x; v[it + 1] = q.y;
#endif
        }
#pragma unroll
        for (u32 it = 0; it < IT; it += 2) {
            const u32 i = 2 * (tid + (it / 2) * THREADS);
            const u32 b = i >> B, g = i & (NPTS - 1);
            tile[g * TP + b] = v[it];
            tile[(g + 1) * TP + b] = v[it + 1];
        }
    }
#ifndef B200ZKP_HOST_EMU
    mbar_wait(bar, 0);
#endif
    NTC_SYNC();

    if (p.a_scale) run_rounds<B, true>(tile, tile, swb, TWP, T_log, TP, p.a_scale);
    else run_rounds<B, false>(tile, tile, swb, TWP, T_log, TP, 0);
    NTC_SYNC();

    NTC_FOR_THREADS(tid) {
        if (NATURAL) {
            // row r of the output = network position g = bitrev_B(r); the T batches of the tile are T consecutive outputs
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                const u32 i = tid + it * THREADS;
                const u32 tau = i & (T - 1), r = i >> T_log;
                const u32 g = bitrev_bits(r, B);
                out[((u64)r << batches_log) + batch0 + tau] = gl::canon_cc(tile[g * TP + tau]);
            }
        } else {
#pragma unroll
            for (u32 it = 0; it < IT; it += 2) {
                const u32 i = 2 * (tid + (it / 2) * THREADS);
                const u32 b = i >> B, g = i & (NPTS - 1);
                const u64 x0 = gl::canon_cc(tile[g * TP + b]), x1 = gl::canon_cc(tile[(g + 1) * TP + b]);
#ifdef B200ZKP_HOST_EMU
                out[((u64)(batch0 + b) << B) + g] = x0; out[((u64)(batch0 + b) << B) + g + 1] = x1;
#else
                *reinterpret_cast<ulonglong2*>(out + ((u64)(batch0 + b) << B) + g) = make_ulonglong2(x0, x1);
#endif
            }
        }
    }
}

#ifndef B200ZKP_HOST_EMU
template <int B, int KIND>
__global__ void __launch_bounds__(THREADS, 4)
ct_pass_kernel(const PassParams p, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ __align__(128) unsigned char ntc_smem_raw[];
    u64* smem = reinterpret_cast<u64*>(ntc_smem_raw);
    if (KIND == KIND_STRIDED) strided_body<B, false>(p, &tm_in, &tm_out, smem, blockIdx.x);
    else if (KIND == KIND_STRIDED_LOOP) strided_body<B, true>(p, &tm_in, &tm_out, smem, blockIdx.x);
    else if (KIND == KIND_FINAL_INPLACE) final_body<B, false>(p, smem, blockIdx.x);
    else final_body<B, true>(p, smem, blockIdx.x);
}

// all-gather + first pass of a partitioned LDE (sharded.inl): tile by tile, the coefficients of every source rank come over
// NVLink from that rank's exchange window into shared memory, from where they are both stored to the local coefficient
// matrix and transformed for this rank's coset blocks.
// Persistent: the launch has a few CTAs per SM (the host picks how many) that walk the tiles with a stride of the grid, and
// the load of a CTA's next tile is in flight while it transforms the current one (two staging slots, one mbarrier each).
// The pass is bound by NVLink, not by the SMs, so it is launched narrow on purpose: the in-place passes of the column
// group gathered before it run beside it on a second stream and fill the rest of every SM.
struct PullMaps { CUtensorMap m[MAX_SRC]; };
template <int B>
struct PullSmem {
    static constexpr u32 NPTS = 1u << B;
    static constexpr u32 sw_words = (u32)MAX_LOOP_BLOCKS * NPTS;
    static constexpr u32 bytes = 4 * TILE * 8 + sw_words * 8 + 32;     // [stage 0][stage 1][work 0][work 1][sw][2 mbarriers]
};

template <int B>
__device__ __forceinline__ void pull_decode(const PassParams& p, u32 bid, u32& col, u32& in_col, u32& q, u32& a, u32& c0) {
    constexpr u32 T_log = 11 - B;
    const u32 vcol = bid % p.ncols, tile_i = bid / p.ncols;
    u32 i;
    col = map_column(p, vcol, q, i);
    in_col = i;
    const u32 tiles_per_a_log = p.C_log - T_log;
    a = tile_i >> tiles_per_a_log;
    c0 = (tile_i & ((1u << tiles_per_a_log) - 1)) << T_log;
}

template <int B>
__global__ void __launch_bounds__(THREADS, 4)
ct_pull_kernel(const PassParams p, const __grid_constant__ PullMaps tm_src, const __grid_constant__ CUtensorMap tm_out,
               const __grid_constant__ CUtensorMap tm_copy, const u32 total) {
    extern __shared__ __align__(128) unsigned char ntc_smem_raw[];
    u64* const smem = reinterpret_cast<u64*>(ntc_smem_raw);
    if (!p.use_tma) {
        // plain-load staging (no tensor maps): the tile body of the coset loop, one tile after the other
        for (u32 bid = blockIdx.x; bid < total; bid += gridDim.x) {
            strided_body<B, true, CUtensorMap, true>(p, tm_src.m, &tm_out, smem, bid, &tm_copy);
            __syncthreads();
        }
        return;
    }
    constexpr u32 NPTS = 1u << B;
    constexpr u32 T_log = 11 - B, T = 1u << T_log;
    u64* const stage0 = smem;
    u64* const work0 = smem + 2 * TILE;
    u64* const sw = smem + 4 * TILE;
    u64* const bar = sw + PullSmem<B>::sw_words;
    if (blockIdx.x >= total) return;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
    __syncthreads();
    u32 col, in_col, q, a, c0;
    if (threadIdx.x == 0) {
        pull_decode<B>(p, blockIdx.x, col, in_col, q, a, c0);
        mbar_expect_tx(bar, TILE * 8);
        tma_load_tile(stage0, &tm_src.m[q], c0, a, 0, in_col, bar);
    }
    u32 it = 0;
    for (u32 bid = blockIdx.x; bid < total; bid += gridDim.x, it++) {
        const u32 slot = it & 1;
        u64* const stage = stage0 + slot * TILE;
        // the other slot is free (every read of the tile before this one was waited for at the end of the last iteration)
        if (threadIdx.x == 0 && bid + gridDim.x < total) {
            pull_decode<B>(p, bid + gridDim.x, col, in_col, q, a, c0);
            mbar_expect_tx(bar + (slot ^ 1), TILE * 8);
            tma_load_tile(stage0 + (slot ^ 1) * TILE, &tm_src.m[q], c0, a, 0, in_col, bar + (slot ^ 1));
        }
        pull_decode<B>(p, bid, col, in_col, q, a, c0);
        const u64 prefix = ((u64)1 << p.S) + a;
        for (u32 i = threadIdx.x; i < p.n_blk * NPTS; i += THREADS) {
            const u32 b = i >> B, e = i & (NPTS - 1);
            if (e) sw[i] = gl::ldg(p.ztab + (u64)b * p.ztab_blk_stride + z_index(prefix, e));
        }
        mbar_wait(bar + slot, (it >> 1) & 1);
        if (threadIdx.x == 0) { tma_store_tile(&tm_copy, c0, a, 0, col, stage); tma_commit(); }
        __syncthreads();
        for (u32 lb = 0; lb < p.n_blk; lb++) {
            u64* const tile = work0 + (lb & 1) * TILE;
            if (lb >= 2) { if (threadIdx.x == 0) tma_wait_read<1>(); __syncthreads(); }
            run_rounds<B, false>(stage, tile, sw + lb * NPTS, 0, T_log, T, 0);
            fence_proxy_async();
            __syncthreads();
            if (threadIdx.x == 0) { tma_store_tile(&tm_out, c0, a, lb, col, tile); tma_commit(); }
        }
        if (threadIdx.x == 0) tma_wait_read<0>();          // stage, work tiles and sw are free for the next tile after this
        __syncthreads();
    }
}

// Z[i], i in [1, n):  l = floor(log2 i), j = i - 2^l:  Z[i] = spow[L - 1 - l] * w^(bitrev_l(j) << (L - 1 - l)) * (l == L-1 ? last_scale : 1)
// spow[e] = s^(2^e) (host computed, L entries per block, in global memory); (lo, hi): two-level powers of w_n (direction specific)
__device__ __forceinline__ u64 z_entry(u64 i, u32 n_log, const u64* __restrict__ spow, const u64* __restrict__ lo,
                                       const u64* __restrict__ hi, u32 lo_bits, u64 last_scale) {
    if (!i) return 0;
    const u32 l = 63u - (u32)__clzll((long long)i);
    const u32 j = (u32)(i - ((u64)1 << l));
    const u64 e = (u64)bitrev_bits(j, l) << (n_log - 1 - l);
    const u64 w = gl::mul(gl::ldg(lo + (e & (((u64)1 << lo_bits) - 1))), gl::ldg(hi + (e >> lo_bits)));
    u64 r = gl::mul(w, gl::ldg(spow + (n_log - 1 - l)));
    if (last_scale && l == n_log - 1) r = gl::mul(r, last_scale);
    return r;
}
// strided passes: out[blk][i] = Z_blk[i] for i < count (the levels before the last pass); grid (ceil(count / 256), n_blk)
__global__ void build_ztab_kernel(u64* __restrict__ out, u64 count, u32 n_log, const u64* __restrict__ spow, const u64* __restrict__ lo,
                                  const u64* __restrict__ hi, u32 lo_bits, u64 last_scale) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[(u64)blockIdx.y * count + i] = z_entry(i, n_log, spow + (u64)blockIdx.y * n_log, lo, hi, lo_bits, last_scale);
}
// final pass of B bits: the tile-ordered image described at FinalSmem; natural: batch b of tile i is row block
// bitrev(i * T + b) (inverse transform).  grid (ceil(words / 256), n_blk), words = (n >> 11) * T * (2^B + 1)
__global__ void build_zfinal_kernel(u64* __restrict__ out, u64 words, u32 n_log, u32 B, u32 natural, const u64* __restrict__ spow,
                                    const u64* __restrict__ lo, const u64* __restrict__ hi, u32 lo_bits, u64 last_scale) {
    const u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= words) return;
    const u32 T_log = 11 - B, TWP = (1u << B) + 1, S = n_log - B;
    const u64 tile_i = x / ((u64)TWP << T_log);
    const u32 rem = (u32)(x - tile_i * ((u64)TWP << T_log));
    const u32 b = rem / TWP, e = rem - b * TWP;
    u64 r = 0;
    if (e >= 1 && e < (1u << B)) {
        const u32 batch = (u32)(tile_i << T_log) + b;
        const u32 a_b = natural ? bitrev_bits(batch, S) : batch;
        r = z_entry(z_index(((u64)1 << S) + a_b, e), n_log, spow + (u64)blockIdx.y * n_log, lo, hi, lo_bits, last_scale);
    }
    out[(u64)blockIdx.y * words + x] = r;
}
#endif  // !B200ZKP_HOST_EMU

}  // namespace ntc
