// Permutation argument of the prover: the Z and partial-product polynomials that prove()
// commits right after the wires.
//
// Replaces plonky2 @ f99ed9c  plonky2/src/plonk/prover.rs `wires_permutation_partial_products_and_zs` (and the helpers
// `quotient_chunk_products`, `partial_products_and_z_gx`), reached from the reference through every prove()
// (/root/reference/src/transaction/circuits/mod.rs:453, src/zkdsa/circuits/mod.rs:326, src/rollup/circuits/mod.rs:1247).
// On the CPU this is one rayon task per row plus a serial running product over the n rows; here it is
//   1. chunk_products_kernel: one thread per (row, challenge): the ceil(R / degree) chunk quotients
//          q_l = prod_{j in chunk l} (w_j + beta k_j x + gamma) / (w_j + beta sigma_j + gamma)
//      (one field inversion per thread: the chunk denominators are inverted together), their running products
//      within the row, and the row product;
//   2. a blocked exclusive prefix product of the row products over the rows (Z(x_i) = prod_{i' < i} row_i', Z(1) = 1):
//      block totals, one CTA scanning the totals, then
//   3. finish_kernel: Z(x_i) and  pp_l = Z(x_i) * (running product l)  written as the columns prove() commits:
//      Z of every challenge first, then the partial products challenge by challenge.
// Wires and sigmas are column-major like every other matrix here (column j at j * stride), values may be non-canonical;
// outputs are canonical.  Exact field arithmetic: any order of multiplications gives plonky2's canonical values.
#pragma once
#include "goldilocks.cuh"

namespace perm {

using gl::u32;
using gl::u64;

static constexpr int MAX_CHUNKS = 32;          // ceil(num_routed_wires / quotient_degree_factor); 10 for the standard config
static constexpr int SCAN_THREADS = 256;

GL_FN u64 pow_u64(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl::mul(r, b);
        b = gl::mul(b, b);
        e >>= 1;
    }
    return r;
}

GL_FN u64 sqr_n(u64 x, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) x = gl::mul(x, x);
    return x;
}

// x^-1 = x^(p - 2),  p - 2 = 2^64 - 2^32 - 1 = (2^32 - 2) * 2^32 + (2^32 - 1).  With t_k = x^(2^k - 1) and
// t_(a+b) = t_a^(2^b) * t_b:  x^(2^32 - 2) = t_31^2,  so  x^(p-2) = (t_31^2)^(2^32) * t_32  (64 squarings, 10 products).
// inverse(0) = 0.
GL_FN u64 inverse(u64 x) {
    const u64 t2 = gl::mul(sqr_n(x, 1), x);
    const u64 t3 = gl::mul(sqr_n(t2, 1), x);
    const u64 t6 = gl::mul(sqr_n(t3, 3), t3);
    const u64 t7 = gl::mul(sqr_n(t6, 1), x);
    const u64 t14 = gl::mul(sqr_n(t7, 7), t7);
    const u64 t15 = gl::mul(sqr_n(t14, 1), x);
    const u64 t30 = gl::mul(sqr_n(t15, 15), t15);
    const u64 t31 = gl::mul(sqr_n(t30, 1), x);
    const u64 t32 = gl::mul(sqr_n(t31, 1), x);
    return gl::mul(sqr_n(t31, 33), t32);
}

struct Params {
    const u64* wires;      // [R][n] column-major
    u64 wires_stride;
    const u64* sigmas;     // [R][n] column-major
    u64 sigmas_stride;
    const u64* k_is;       // [R] canonical (device)
    const u64* betas;      // [C] canonical (device)
    const u64* gammas;     // [C]
    u64 omega;             // w_n
    u64 n;
    u32 R, degree, chunks, C;
};

// running[(c * chunks + l) * n + i] = q_0 * ... * q_l of row i, challenge c  (l = chunks - 1: the row product)
GL_FN void row_chunk_products(const Params& p, u64 i, u32 c, u64* __restrict__ running) {
    const u64 beta = p.betas[c], gamma = p.gammas[c];
    const u64 bx = gl::mul(beta, pow_u64(p.omega, i));
    u64 num[MAX_CHUNKS], den[MAX_CHUNKS];
#pragma unroll 1
    for (u32 l = 0; l < p.chunks; l++) {
        u64 nn = 1, dd = 1;
        const u32 j1 = (l + 1) * p.degree < p.R ? (l + 1) * p.degree : p.R;
#pragma unroll 1
        for (u32 j = l * p.degree; j < j1; j++) {
            const u64 w = gl::canon(p.wires[(u64)j * p.wires_stride + i]);
            const u64 s = gl::canon(p.sigmas[(u64)j * p.sigmas_stride + i]);
            const u64 wg = gl::add(w, gamma);
            nn = gl::mul(nn, gl::add(wg, gl::mul(bx, p.k_is[j])));
            dd = gl::mul(dd, gl::add(wg, gl::mul(beta, s)));
        }
        num[l] = nn;
        den[l] = dd;
    }
    // invert the chunk denominators together: prefix products, one inversion, walk back
    u64 pre[MAX_CHUNKS];
    u64 acc = 1;
#pragma unroll 1
    for (u32 l = 0; l < p.chunks; l++) { pre[l] = acc; acc = gl::mul(acc, den[l]); }
    u64 inv_all = inverse(acc);
#pragma unroll 1
    for (u32 l = p.chunks; l-- > 0;) {
        const u64 inv_l = gl::mul(inv_all, pre[l]);
        inv_all = gl::mul(inv_all, den[l]);
        num[l] = gl::mul(num[l], inv_l);            // q_l
    }
    acc = 1;
#pragma unroll 1
    for (u32 l = 0; l < p.chunks; l++) {
        acc = gl::mul(acc, num[l]);
        running[((u64)c * p.chunks + l) * p.n + i] = acc;
    }
}

#ifndef B200ZKP_HOST_EMU
__global__ void __launch_bounds__(128)
chunk_products_kernel(Params p, u64* __restrict__ running) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.n) row_chunk_products(p, i, blockIdx.y, running);
}

// product of the row products of one block of SCAN_THREADS * per rows -> totals[c * n_blocks + b]
__global__ void __launch_bounds__(SCAN_THREADS)
block_totals_kernel(const u64* __restrict__ running, u64 n, u32 chunks, u32 per, u64* __restrict__ totals) {
    __shared__ u64 red[SCAN_THREADS];
    const u32 t = threadIdx.x, c = blockIdx.y;
    const u64* row = running + ((u64)c * chunks + (chunks - 1)) * n;
    const u64 base = ((u64)blockIdx.x * SCAN_THREADS + t) * per;
    u64 acc = 1;
    for (u32 e = 0; e < per; e++) if (base + e < n) acc = gl::mul(acc, row[base + e]);
    red[t] = acc;
    __syncthreads();
    for (u32 s = SCAN_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] = gl::mul(red[t], red[t + s]);
        __syncthreads();
    }
    if (t == 0) totals[(u64)c * gridDim.x + blockIdx.x] = red[0];
}

// exclusive prefix product of the block totals of every challenge, in place (one CTA per challenge, serial chunks per thread)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_totals_kernel(u64* __restrict__ totals, u32 n_blocks) {
    __shared__ u64 sh[SCAN_THREADS];
    const u32 t = threadIdx.x;
    u64* tot = totals + (u64)blockIdx.x * n_blocks;
    const u32 per = (n_blocks + SCAN_THREADS - 1) / SCAN_THREADS;
    const u32 b0 = t * per;
    u64 acc = 1;
    for (u32 e = 0; e < per; e++) if (b0 + e < n_blocks) acc = gl::mul(acc, tot[b0 + e]);
    sh[t] = acc;
    __syncthreads();
    for (u32 s = 1; s < SCAN_THREADS; s <<= 1) {        // Hillis-Steele inclusive scan
        u64 v = (t >= s) ? gl::mul(sh[t - s], sh[t]) : sh[t];
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    u64 carry = t ? sh[t - 1] : 1;
    for (u32 e = 0; e < per; e++) {
        if (b0 + e < n_blocks) {
            const u64 v = tot[b0 + e];
            tot[b0 + e] = carry;
            carry = gl::mul(carry, v);
        }
    }
}

// Z(x_i) from the block carry and the rows before i inside the block, then the partial products of row i.
// out: column c = Z of challenge c; column C + c * num_prods + l = partial product l of challenge c
__global__ void __launch_bounds__(SCAN_THREADS)
finish_kernel(const u64* __restrict__ running, const u64* __restrict__ carries, u64 n, u32 chunks, u32 per, u32 C,
              u64* __restrict__ out, u64 out_stride) {
    __shared__ u64 sh[SCAN_THREADS];
    const u32 t = threadIdx.x, c = blockIdx.y;
    const u64* row = running + ((u64)c * chunks + (chunks - 1)) * n;
    const u64 base = ((u64)blockIdx.x * SCAN_THREADS + t) * per;
    u64 acc = 1;
    for (u32 e = 0; e < per; e++) if (base + e < n) acc = gl::mul(acc, row[base + e]);
    sh[t] = acc;
    __syncthreads();
    for (u32 s = 1; s < SCAN_THREADS; s <<= 1) {
        u64 v = (t >= s) ? gl::mul(sh[t - s], sh[t]) : sh[t];
        __syncthreads();
        sh[t] = v;
        __syncthreads();
    }
    u64 z = gl::mul(carries[(u64)c * gridDim.x + blockIdx.x], t ? sh[t - 1] : 1);
    const u32 num_prods = chunks - 1;
    for (u32 e = 0; e < per; e++) {
        const u64 i = base + e;
        if (i >= n) break;
        out[(u64)c * out_stride + i] = z;
        for (u32 l = 0; l < num_prods; l++)
            out[((u64)C + (u64)c * num_prods + l) * out_stride + i] = gl::mul(z, running[((u64)c * chunks + l) * n + i]);
        z = gl::mul(z, row[i]);
    }
}

#endif  // !B200ZKP_HOST_EMU

}  // namespace perm
