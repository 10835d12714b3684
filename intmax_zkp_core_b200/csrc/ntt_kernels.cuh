// Goldilocks NTT passes for sm_100a: inverse NTT (values -> coefficients) and the coset
// low-degree extension written directly in bit-reversed (leaf) order.
//
// Replaces plonky2_field `fft_classic` / `ifft_with_options` / `coset_fft_with_options` / `lde`
// (plonky2 @ f99ed9c, field/src/fft.rs, field/src/polynomial/mod.rs) and, for the LDE, also
// plonky2 `transpose` + `reverse_index_bits_in_place` (plonky2/src/util) — SURVEY.md rows A2, A4, A5,
// A12.  Reached from the reference through PolynomialBatch::from_values / from_coeffs inside every
// prove()/build(), e.g. /root/reference/src/transaction/circuits/mod.rs:158,453.
//
// A transform of 2^L points is split into P = ceil(L/8) passes, one pass for L <= 10 (Cooley-Tukey / "4-step"):
//   pass p views a column as [A][2^B][C]; a CTA stages a tile of 2^B x T elements (T batches of the
//   contiguous inner dimension, so every global access is a run of >= 64 B) in shared memory, runs the
//   2^B-point decimation-in-frequency sub-transform with radix-8 butterflies held in registers (one
//   shared-memory round trip per three stages), multiplies by the inter-pass twiddle w_M^(c*k1) and
//   writes the tile back.
//   * bit-reversed output (LDE): every pass writes in place -> the result is the plain DIF order,
//     which IS plonky2's leaf order inside one coset, so no transpose / bit-reversal pass exists.
//   * natural output (iNTT): non-final passes write k1 in natural position, the final pass scatters
//     with the digit-reversed batch index (runs of T contiguous elements).
// All values are canonical (< p) inside the passes.
#pragma once
#include "goldilocks.cuh"

namespace ntt {

using gl::u32;
using gl::u64;

static constexpr int TILE_ELEMS = 2048;  // elements staged per CTA (passes of up to 8 bits)
static constexpr int THREADS = 256;      // 8 elements per thread
static constexpr int MAX_PASS_BITS = 10; // 9- and 10-bit passes stage 4096 elements with 512 threads (>= 4 batches per tile)
static constexpr int MAX_PASSES = 4;
#ifdef __CUDACC__
#define NTT_CONSTEXPR_FN __host__ __device__ constexpr
#else
#define NTT_CONSTEXPR_FN constexpr
#endif
NTT_CONSTEXPR_FN int tile_elems_for(int B) { return B >= 9 ? 2 * TILE_ELEMS : TILE_ELEMS; }
NTT_CONSTEXPR_FN int threads_for(int B) { return B >= 9 ? 2 * THREADS : THREADS; }

enum OutMode : u32 {
    OUT_INPLACE_NATURAL = 0,  // non-final pass, element k1 stored at (a, k1, c)
    OUT_INPLACE_BITREV = 1,   // any pass, element k1 stored at (a, bitrev(k1), c)
    OUT_FINAL_NATURAL = 2,    // final pass, digit-reversed scatter to natural order
};

struct PassParams {
    const u64* in;
    u64* out;
    u64 in_col_stride, out_col_stride;
    u64 in_blk_stride, out_blk_stride;   // element offsets per blockIdx.y (coset block of an LDE)
    u32 ncols;
    u32 n_log;     // log2 of the column transform size
    u32 C_log;     // log2 of the inner (contiguous) extent below this pass's dimension
    u32 out_mode;
    const u64* wtab;    // w_{2^B}^e for e < 2^(B-1), direction specific
    const u64* scale;   // optional input scaling, direct table [blk][n]: scale[blk*scale_blk_stride + i] (coset shift^i)
    u64 scale_blk_stride;
    const u64* twimg;   // optional inter-pass twiddle image, 2^(B+C_log) entries laid out like the output block:
                        // twimg[pos*C + c] = w_M^(c*k1) (k1 = pos, or bitrev(pos) for OUT_INPLACE_BITREV)
    u64 out_scale;      // 0: none; else multiply every output (n^-1 of a single-pass inverse transform)
    u32 canon_in;       // inputs may be non-canonical
    u32 n_digits;       // OUT_FINAL_NATURAL: bit widths of the earlier passes, first pass first
    u32 digits[MAX_PASSES];
};

GL_FN u32 bitrev(u32 x, u32 bits) { return bits ? (gl::brev32(x) >> (32 - bits)) : 0; }

// The pass body is written once and compiled two ways: on the device one CUDA thread runs each
// NTT_FOR_THREADS body and NTT_SYNC is __syncthreads(); under B200ZKP_HOST_EMU (tests only) the same
// text steps all 256 "threads" of a CTA in a loop so the index logic can be checked without a GPU.
#ifdef B200ZKP_HOST_EMU
#define NTT_FOR_THREADS(tid) for (u32 tid = 0; tid < THREADS; tid++)
#define NTT_SYNC() do {} while (0)
#define NTT_SHARED static thread_local
#else
#define NTT_FOR_THREADS(tid) for (u32 tid = threadIdx.x, once__ = 1; once__; once__ = 0)
#define NTT_SYNC() __syncthreads()
#define NTT_SHARED __shared__
#endif

// in-register DIF over 2^A elements x[j] <-> g = gbase + j*st ; stages sigma0 .. sigma0+A-1 of a 2^B transform
// LAST: the round that ends the 2^B-point transform (st == 1, ul == 0): the twiddle exponent is jl << (sigma0 + u), a
// compile-time zero for jl == 0 — 7 of the 12 butterflies of a radix-8 group multiply by 1 and skip the product.
template <int A, bool LAST>
GL_FN void dif_group(u64 (&x)[1 << A], const u64* __restrict__ wt, u32 ul, u32 st, u32 sigma0) {
#pragma unroll
    for (int u = 0; u < A; u++) {
        const int half = (1 << A) >> (u + 1);
#pragma unroll
        for (int j = 0; j < (1 << A); j++) {
            if (j & half) continue;
            u32 jl = j & (half - 1);
            u64 a = x[j], b = x[j + half];
            x[j] = gl::add(a, b);
            if (LAST && jl == 0) {
                x[j + half] = gl::sub(a, b);
            } else {
                u32 e = LAST ? (jl << (sigma0 + u)) : ((jl * st + ul) << (sigma0 + u));
                x[j + half] = gl::mul(gl::sub(a, b), wt[e]);
            }
        }
    }
}

template <int A, int NTHREADS, bool LAST = false>
GL_FN void run_round(u64* __restrict__ tile, const u64* __restrict__ wt, u32 B, u32 T, u32 TP, u32 sigma0,
                     u32 tid) {
    const u32 st_log = LAST ? 0u : B - sigma0 - A;
    const u32 st = 1u << st_log;
    constexpr int UNITS_PER_THREAD = 8 >> A;
#pragma unroll
    for (int r = 0; r < UNITS_PER_THREAD; r++) {
        u32 U = tid + r * NTHREADS;
        u32 t = U % T, w = U / T;
        u32 ul = w & (st - 1), uh = w >> st_log;
        u32 gbase = (uh << (st_log + A)) + ul;
        u64 x[1 << A];
#pragma unroll
        for (int j = 0; j < (1 << A); j++) x[j] = tile[(gbase + j * st) * TP + t];
        dif_group<A, LAST>(x, wt, ul, st, sigma0);
#pragma unroll
        for (int j = 0; j < (1 << A); j++) tile[(gbase + j * st) * TP + t] = x[j];
    }
}

GL_FN u64 two_level(const u64* __restrict__ lo, const u64* __restrict__ hi, u32 lo_bits, u64 e) {
    u64 a = gl::ldg(lo + (e & (((u64)1 << lo_bits) - 1)));
    u64 b = gl::ldg(hi + (e >> lo_bits));
    return gl::mul(a, b);
}

// digit reversal of a batch index: in-place position digits (d1 most significant) <-> natural
// batch index beta' = d1 + 2^B1 * (d2 + 2^B2 * ...)
GL_FN u64 digit_reverse(u64 beta_nat, const PassParams& p) {
    u64 pos = 0;
#pragma unroll
    for (u32 i = 0; i < (u32)MAX_PASSES; i++) {   // static indices keep the parameter block out of local memory
        if (i < p.n_digits) {
            u32 w = p.digits[i];
            pos = (pos << w) | (beta_nat & (((u64)1 << w) - 1));
            beta_nat >>= w;
        }
    }
    return pos;
}

// Kernel flavours (compile time, so each instantiation carries only its own load/store path and stays small in
// the instruction cache): non-final passes in natural / bit-reversed position, final pass in place (LDE) or
// scattered to natural order (inverse transform).
enum PassMode : int { MODE_MID_NATURAL = 0, MODE_MID_BITREV = 1, MODE_FINAL_BITREV = 2, MODE_FINAL_NATURAL = 3 };
static inline int pass_mode(const PassParams& p) {   // host side (launcher / emulation harness)
    if (p.out_mode == OUT_FINAL_NATURAL) return MODE_FINAL_NATURAL;
    if (p.out_mode == OUT_INPLACE_NATURAL) return MODE_MID_NATURAL;
    return p.C_log == 0 ? MODE_FINAL_BITREV : MODE_MID_BITREV;
}

// B = bits of this pass (1..8, compile time so the rounds fully unroll); block = CTA index, blk = coset block
template <int B, int MODE>
GL_FN void pass_body(const PassParams& p, u32 block, u32 blk) {
    constexpr int TILE_ELEMS = tile_elems_for(B);        // (shadow the 8-bit defaults of the namespace)
    constexpr int THREADS = threads_for(B);
    constexpr u32 T = TILE_ELEMS >> B;
    constexpr u32 TP = T + 1;
    constexpr u32 NPTS = 1u << B;
    constexpr u32 IT = TILE_ELEMS / THREADS;             // elements per thread
    constexpr bool kOwnBatch = (T <= (u32)THREADS);      // strided tiles: a thread keeps one batch for all its elements
    constexpr u32 GSTEP = kOwnBatch ? (THREADS / T) : 1; // row step between a thread's consecutive elements
    NTT_SHARED u64 tile[NPTS * TP];
    NTT_SHARED u64 wt[(NPTS / 2) ? (NPTS / 2) : 1];

    const u32 batches_log = p.n_log - B;                 // batches per column
    const u64 total_batches = (u64)p.ncols << batches_log;
    const u64 batch_mask = ((u64)1 << batches_log) - 1;
    const u64 tile_b0 = (u64)block * T;
    constexpr bool final_pass = (MODE == MODE_FINAL_BITREV || MODE == MODE_FINAL_NATURAL);
    constexpr u32 out_mode = MODE == MODE_MID_NATURAL ? OUT_INPLACE_NATURAL : (MODE == MODE_FINAL_NATURAL ? OUT_FINAL_NATURAL : OUT_INPLACE_BITREV);
    // final in-place tiles are one contiguous run of TILE_ELEMS elements unless a tile straddles two columns
    const bool whole_tile_in_column = (((u64)1 << batches_log) >= T) && (tile_b0 + T <= total_batches);
    const u32 C_mask = (u32)(((u64)1 << p.C_log) - 1);
    const u64* __restrict__ in = p.in + (u64)blk * p.in_blk_stride;
    u64* __restrict__ out = p.out + (u64)blk * p.out_blk_stride;
    const u64* __restrict__ scale = p.scale ? p.scale + (u64)blk * p.scale_blk_stride : nullptr;

    // ---- load (all of a thread's loads are issued before the first use)
    NTT_FOR_THREADS(tid) {
        for (u32 i = tid; i < NPTS / 2; i += THREADS) wt[i] = p.wtab[i];
        u64 v[IT];
        u32 sm[IT];
        if (!final_pass && kOwnBatch) {
            const u32 b = tid % T, g0 = tid / T;
            const u64 bg = tile_b0 + b;
            const bool valid = bg < total_batches;
            u64 base = 0;
            u32 icol0 = 0;
            if (valid) {
                u64 col = bg >> batches_log;
                u32 beta = (u32)(bg & batch_mask);
                u32 a = beta >> p.C_log, c = beta & C_mask;
                icol0 = (a << (B + p.C_log)) + c;
                base = col * p.in_col_stride + icol0;
            }
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                u32 g = g0 + it * GSTEP;
                sm[it] = g * TP + b;
                v[it] = valid ? in[base + ((u64)g << p.C_log)] : 0;
            }
            if (scale && valid) {
                u64 f[IT];
#pragma unroll
                for (u32 it = 0; it < IT; it++) f[it] = scale[icol0 + ((g0 + it * GSTEP) << p.C_log)];
#pragma unroll
                for (u32 it = 0; it < IT; it++) v[it] = gl::mul(p.canon_in ? gl::canon(v[it]) : v[it], f[it]);
            } else if (p.canon_in) {
#pragma unroll
                for (u32 it = 0; it < IT; it++) v[it] = gl::canon(v[it]);
            }
        } else if (MODE == MODE_FINAL_BITREV && whole_tile_in_column) {
            // the tile is the contiguous run [base, base + TILE_ELEMS) of one column
            const u64 base = (tile_b0 >> batches_log) * p.in_col_stride + ((tile_b0 & batch_mask) << B);
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                u32 idx = tid + it * THREADS;
                sm[it] = (idx & (NPTS - 1)) * TP + (idx >> B);
                v[it] = in[base + idx];
            }
            if (scale) {
                const u64 sbase = (tile_b0 & batch_mask) << B;
#pragma unroll
                for (u32 it = 0; it < IT; it++)
                    v[it] = gl::mul(p.canon_in ? gl::canon(v[it]) : v[it], scale[sbase + tid + it * THREADS]);
            } else if (p.canon_in) {
#pragma unroll
                for (u32 it = 0; it < IT; it++) v[it] = gl::canon(v[it]);
            }
        } else {
            // generic path: final natural-order pass, tiles straddling columns, very small transforms
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                u32 idx = tid + it * THREADS;
                u32 b, g;
                if (final_pass) { b = idx >> B; g = idx & (NPTS - 1); }
                else { b = idx % T; g = idx / T; }
                sm[it] = g * TP + b;
                u64 bg = tile_b0 + b;
                u64 x = 0;
                if (bg < total_batches) {
                    u64 col = bg >> batches_log;
                    u64 beta = bg & batch_mask;
                    if (out_mode == OUT_FINAL_NATURAL) beta = digit_reverse(beta, p);
                    u64 a = beta >> p.C_log, c = beta & C_mask;
                    u64 i_col = (a << (B + p.C_log)) + ((u64)g << p.C_log) + c;
                    x = in[col * p.in_col_stride + i_col];
                    if (p.canon_in) x = gl::canon(x);
                    if (scale) x = gl::mul(x, scale[i_col]);
                }
                v[it] = x;
            }
        }
#pragma unroll
        for (u32 it = 0; it < IT; it++) tile[sm[it]] = v[it];
    }
    NTT_SYNC();

    // ---- 2^B-point DIF, radix-8 in registers
    {
        u32 sigma = 0;
        // the short round (B mod 3 levels) goes first so that the closing round is a full radix-8 group with unit stride:
        // 7 of its 12 twiddles are 1 and known at compile time
        if (B > 3 && B % 3 == 2) { NTT_FOR_THREADS(tid) { run_round<2, THREADS>(tile, wt, B, T, TP, sigma, tid); } sigma += 2; NTT_SYNC(); }
        if (B > 3 && B % 3 == 1) { NTT_FOR_THREADS(tid) { run_round<1, THREADS>(tile, wt, B, T, TP, sigma, tid); } sigma += 1; NTT_SYNC(); }
#pragma unroll 1   // one copy of the radix-8 round in the instruction cache; stride and twiddle step are run-time
        for (int r = 0; r < B / 3 - 1; r++) {
            NTT_FOR_THREADS(tid) { run_round<3, THREADS>(tile, wt, B, T, TP, sigma, tid); }
            sigma += 3;
            NTT_SYNC();
        }
        if (B >= 3) { NTT_FOR_THREADS(tid) { run_round<3, THREADS, true>(tile, wt, B, T, TP, sigma, tid); } NTT_SYNC(); }
        if (B == 2) { NTT_FOR_THREADS(tid) { run_round<2, THREADS, true>(tile, wt, B, T, TP, sigma, tid); } NTT_SYNC(); }
        if (B == 1) { NTT_FOR_THREADS(tid) { run_round<1, THREADS, true>(tile, wt, B, T, TP, sigma, tid); } NTT_SYNC(); }
    }

    // ---- twiddle + store
    NTT_FOR_THREADS(tid) {
        constexpr bool contiguous = (MODE == MODE_FINAL_BITREV);
        if (contiguous && whole_tile_in_column) {
            const u64 base = (tile_b0 >> batches_log) * p.out_col_stride + ((tile_b0 & batch_mask) << B);
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                u32 idx = tid + it * THREADS;
                u64 x = tile[(idx & (NPTS - 1)) * TP + (idx >> B)];
                out[base + idx] = p.out_scale ? gl::mul(x, p.out_scale) : x;
            }
        } else if (!contiguous && kOwnBatch) {
            // row = k1 (natural modes) or g~ (bit-reversed mode); the thread keeps batch b
            const u32 b = tid % T, r0 = tid / T;
            const u64 bg = tile_b0 + b;
            if (bg < total_batches) {
                u64 col = bg >> batches_log;
                u32 beta = (u32)(bg & batch_mask);
                u64 v[IT];
                u32 pos[IT], kk[IT];
#pragma unroll
                for (u32 it = 0; it < IT; it++) {
                    u32 row = r0 + it * GSTEP;
                    u32 gt, k1;
                    if (out_mode == OUT_INPLACE_BITREV) { gt = row; k1 = bitrev(row, B); }
                    else { k1 = row; gt = bitrev(row, B); }
                    v[it] = tile[gt * TP + b];
                    pos[it] = (out_mode == OUT_INPLACE_BITREV) ? gt : k1;
                    kk[it] = k1;
                }
                if (out_mode == OUT_FINAL_NATURAL) {
                    u64 obase = col * p.out_col_stride + beta;
#pragma unroll
                    for (u32 it = 0; it < IT; it++) {
                        u64 x = p.out_scale ? gl::mul(v[it], p.out_scale) : v[it];
                        out[obase + ((u64)kk[it] << batches_log)] = x;
                    }
                } else {
                    u32 a = beta >> p.C_log, c = beta & C_mask;
                    u64 obase = col * p.out_col_stride + ((u64)a << (B + p.C_log)) + c;
                    if (p.twimg) {
                        u64 f[IT];
#pragma unroll
                        for (u32 it = 0; it < IT; it++) f[it] = gl::ldg(p.twimg + ((u64)pos[it] << p.C_log) + c);
#pragma unroll
                        for (u32 it = 0; it < IT; it++) v[it] = gl::mul(v[it], f[it]);
                    }
#pragma unroll
                    for (u32 it = 0; it < IT; it++) {
                        u64 x = p.out_scale ? gl::mul(v[it], p.out_scale) : v[it];
                        out[obase + ((u64)pos[it] << p.C_log)] = x;
                    }
                }
            }
        } else {
#pragma unroll
            for (u32 it = 0; it < IT; it++) {
                u32 idx = tid + it * THREADS;
                u32 b, row;
                if (contiguous) { b = idx >> B; row = idx & (NPTS - 1); }
                else { b = idx % T; row = idx / T; }
                u64 bg = tile_b0 + b;
                if (bg >= total_batches) continue;
                u32 gt, k1;
                if (out_mode == OUT_INPLACE_BITREV) { gt = row; k1 = bitrev(row, B); }
                else { k1 = row; gt = bitrev(row, B); }
                u64 x = tile[gt * TP + b];
                u64 col = bg >> batches_log;
                u64 beta = bg & batch_mask;
                u64 o_col;
                if (out_mode == OUT_FINAL_NATURAL) {
                    o_col = beta + ((u64)k1 << batches_log);
                } else {
                    u64 a = beta >> p.C_log, c = beta & C_mask;
                    u32 pos = (out_mode == OUT_INPLACE_BITREV) ? gt : k1;
                    o_col = (a << (B + p.C_log)) + ((u64)pos << p.C_log) + c;
                    if (p.twimg) x = gl::mul(x, gl::ldg(p.twimg + ((u64)pos << p.C_log) + c));
                }
                if (p.out_scale) x = gl::mul(x, p.out_scale);
                out[col * p.out_col_stride + o_col] = x;
            }
        }
    }
}

#ifndef B200ZKP_HOST_EMU
template <int B, int MODE>
__global__ void __launch_bounds__(threads_for(B), B >= 9 ? 2 : 4) ntt_pass_kernel(PassParams p) { pass_body<B, MODE>(p, blockIdx.x, blockIdx.y); }

// table builders (run once per (n_log, direction, rate_bits) and cached by the context)
// out[i] = base^i for i < count, from the two-level power tables of `base`
__global__ void build_powers_kernel(u64* __restrict__ out, u64 count, const u64* __restrict__ lo, const u64* __restrict__ hi,
                                    u32 lo_bits) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = two_level(lo, hi, lo_bits, i);
}
// out[pos*C + c] = f * w_n^((c * k1) << shift), k1 = pos or bitrev(pos, B); (lo, hi) are the two-level powers of w_n
__global__ void build_twiddle_image_kernel(u64* __restrict__ out, u32 B, u32 C_log, u32 shift, u32 bitrev_pos, u64 f,
                                           const u64* __restrict__ lo, const u64* __restrict__ hi, u32 lo_bits) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((u64)1 << (B + C_log))) return;
    u32 pos = (u32)(i >> C_log);
    u64 c = i & (((u64)1 << C_log) - 1);
    u32 k1 = bitrev_pos ? bitrev(pos, B) : pos;
    u64 w = two_level(lo, hi, lo_bits, (c * k1) << shift);
    out[i] = f ? gl::mul(w, f) : w;
}

// dst[c][bitrev(i)] = canon(src[c][i]) : salt columns enter the leaves in bit-reversed row order (A4/A5)
__global__ void bitrev_copy_kernel(const u64* __restrict__ src, u64* __restrict__ dst, u32 n_log, u32 ncols,
                                   u64 src_stride, u64 dst_stride) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 n = (u64)1 << n_log;
    if (g >= n * ncols) return;
    u64 c = g >> n_log, i = g & (n - 1);
    u64 j = n_log ? (gl::brev64(i) >> (64 - n_log)) : 0;
    dst[c * dst_stride + j] = gl::canon(src[c * src_stride + i]);
}

// data[c][j] *= pw[j]  (coefficient scaling of a coset inverse transform); grid (ceil(n / 256), k)
__global__ void scale_columns_kernel(u64* __restrict__ data, u64 stride, u64 n, const u64* __restrict__ pw) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { u64* p = data + (u64)blockIdx.y * stride + j; *p = gl::mul(*p, gl::ldg(pw + j)); }
}

__global__ void canon_copy_kernel(const u64* __restrict__ src, u64* __restrict__ dst, u64 count) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < count) dst[g] = gl::canon(src[g]);
}

#endif  // !B200ZKP_HOST_EMU

}  // namespace ntt
