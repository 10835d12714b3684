// Openings (first piece): evaluate every polynomial of a batch at one point of the
// quadratic extension F_p[X]/(X^2 - 7).
//
// Replaces the evaluation loop of plonky2 `OpeningSet::new` / `eval_commitment`:
//     c.polynomials.par_iter().map(|p| p.to_extension().eval(z))          (plonky2/src/plonk/proof.rs, @ f99ed9c)
// with `QuadraticExtension<GoldilocksField>` arithmetic (field/src/extension/quadratic.rs, W = 7), reached from the
// reference through every prove() (e.g. /root/reference/src/rollup/circuits/mod.rs:1247).  The coefficients are already
// in HBM after the commitment, so the prover no longer has to copy 8nk bytes back just to open them.
//
// p(z) = sum_t z^t * Q_t(z^256),  Q_t(y) = sum_j c[256 j + t] y^j : thread t of a CTA runs Horner in y over its strided
// (coalesced) coefficients; columns are cut into segments so the grid fills the GPU; a second tiny kernel adds the
// segment partials  sum_s z^(s * seg_len) * P_s.
#pragma once
#include "goldilocks.cuh"

namespace evalk {

using gl::u32;
using gl::u64;

struct Ext2 { u64 a, b; };   // a + b X, X^2 = 7, both canonical

GL_FN Ext2 ext_mul(Ext2 x, Ext2 y) {
    u64 t0 = gl::mul(x.a, y.a), t1 = gl::mul(x.b, y.b);
    u64 t2 = gl::mul(x.a, y.b), t3 = gl::mul(x.b, y.a);
    Ext2 r;
    r.a = gl::add(t0, gl::mul(t1, 7));
    r.b = gl::add(t2, t3);
    return r;
}
GL_FN Ext2 ext_add(Ext2 x, Ext2 y) { Ext2 r; r.a = gl::add(x.a, y.a); r.b = gl::add(x.b, y.b); return r; }
GL_FN Ext2 ext_pow(Ext2 x, u64 e) {
    Ext2 r; r.a = 1; r.b = 0;
    while (e) { if (e & 1) r = ext_mul(r, x); x = ext_mul(x, x); e >>= 1; }
    return r;
}

static constexpr int EVAL_THREADS = 256;

#ifndef B200ZKP_HOST_EMU
// grid (segments, columns); partial[col][seg] = sum_{i in segment} c[i] z^(i - seg_start)
__global__ void __launch_bounds__(EVAL_THREADS)
eval_segments_kernel(const u64* __restrict__ coeffs, u64 col_stride, u64 n, u64 seg_len, Ext2 z, Ext2* __restrict__ partial) {
    __shared__ Ext2 red[EVAL_THREADS];
    const u32 t = threadIdx.x;
    const u64 seg = blockIdx.x, col = blockIdx.y;
    const u64 i0 = seg * seg_len;
    const u64 i1 = (i0 + seg_len < n) ? i0 + seg_len : n;
    const u64* c = coeffs + col * col_stride;
    const Ext2 y = ext_pow(z, EVAL_THREADS);
    // this thread's coefficients: i0 + t, i0 + t + 256, ...; Horner from the top
    Ext2 acc; acc.a = 0; acc.b = 0;
    if (i0 + t < i1) {
        u64 cnt = (i1 - i0 - t + EVAL_THREADS - 1) / EVAL_THREADS;
        for (u64 j = cnt; j-- > 0;) {
            acc = ext_mul(acc, y);
            acc.a = gl::add(acc.a, gl::canon(c[i0 + t + j * EVAL_THREADS]));
        }
        acc = ext_mul(acc, ext_pow(z, t));
    }
    red[t] = acc;
    __syncthreads();
    for (u32 s = EVAL_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] = ext_add(red[t], red[t + s]);
        __syncthreads();
    }
    if (t == 0) partial[col * gridDim.x + seg] = red[0];
}

// out[col] = sum_s z^(s * seg_len) * partial[col][s]
__global__ void eval_combine_kernel(const Ext2* __restrict__ partial, u32 n_seg, u64 seg_len, u32 k, Ext2 z, Ext2* __restrict__ out) {
    u32 col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= k) return;
    Ext2 step = ext_pow(z, seg_len);
    Ext2 acc; acc.a = 0; acc.b = 0;
    for (u32 s = n_seg; s-- > 0;) acc = ext_add(ext_mul(acc, step), partial[(u64)col * n_seg + s]);
    out[col] = acc;
}
#endif

}  // namespace evalk
