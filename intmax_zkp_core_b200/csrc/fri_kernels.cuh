// Opening proof (SURVEY.md section 8f, rows N2 + N3): the data-parallel steps of plonky2's
//   PolynomialBatch::prove_openings      plonky2/src/fri/oracle.rs      (@ f99ed9c, un-vendored dependency)
//   fri_committed_trees / fri_proof_of_work / fri_prover_query_round    plonky2/src/fri/prover.rs
//   ReducingFactor::reduce_polys_base / shift_poly                      plonky2/src/util/reducing.rs
//   PolynomialCoeffs::divide_by_linear                                  field/src/polynomial/division.rs
// reached from the reference through every prove() (e.g. /root/reference/src/rollup/circuits/mod.rs:1247).
//
// Extension-field polynomials live in HBM as two planes  [2][len]  (plane 0: constant parts, plane 1: X parts of
// F_p[X]/(X^2 - 7)), so the base-field transform kernels run on them as two columns; the Merkle leaves of a FRI layer are
// the interleaved rows [len / arity][2 * arity] plonky2's `flatten` produces.
#pragma once
#include "eval_kernels.cuh"
#include "poseidon.cuh"

namespace frik {

using evalk::Ext2;
using evalk::ext_add;
using evalk::ext_mul;
using evalk::ext_pow;
using gl::u32;
using gl::u64;

GL_FN Ext2 ext_zero() { Ext2 r; r.a = 0; r.b = 0; return r; }
GL_FN Ext2 ext_scale(Ext2 x, u64 c) { Ext2 r; r.a = gl::mul(x.a, c); r.b = gl::mul(x.b, c); return r; }

static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_PER_THREAD = 4;
static constexpr int SCAN_SEG = SCAN_THREADS * SCAN_PER_THREAD;

#ifndef B200ZKP_HOST_EMU
// pw[i] = base^i
__global__ void ext_powers_kernel(Ext2 base, u32 count, Ext2* __restrict__ pw) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) pw[i] = ext_pow(base, i);
}

// ReducingFactor::reduce_polys_base: out[j] = sum_i alpha^i * cols[i][j]   (base-field polynomials, extension scalar).
// One thread per coefficient index, coalesced over j; the pointer and power tables are warp-uniform loads.
__global__ void __launch_bounds__(256)
reduce_polys_base_kernel(const u64* const* __restrict__ cols, u32 k, u64 n, const Ext2* __restrict__ pw,
                         u64* __restrict__ out_a, u64* __restrict__ out_b) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    u64 a = 0, b = 0;
    for (u32 i = 0; i < k; i++) {
        u64 c = gl::ldg(cols[i] + j);
        Ext2 p = pw[i];
        a = gl::add(a, gl::mul(p.a, c));
        b = gl::add(b, gl::mul(p.b, c));
    }
    out_a[j] = a;
    out_b[j] = b;
}

// Segment totals of the suffix Horner scan: tot[s] = sum_{i in segment s} c[i] z^(i - start_s), segments of SCAN_SEG.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_totals_kernel(const u64* __restrict__ in_a, const u64* __restrict__ in_b, u64 n, Ext2 z, Ext2* __restrict__ tot) {
    __shared__ Ext2 red[SCAN_THREADS];
    const u32 t = threadIdx.x;
    const u64 base = (u64)blockIdx.x * SCAN_SEG + (u64)t * SCAN_PER_THREAD;
    Ext2 h = ext_zero();
#pragma unroll
    for (int e = SCAN_PER_THREAD - 1; e >= 0; e--) {
        Ext2 c = ext_zero();
        if (base + e < n) { c.a = gl::canon(in_a[base + e]); c.b = gl::canon(in_b[base + e]); }
        h = ext_add(ext_mul(h, z), c);
    }
    red[t] = ext_mul(h, ext_pow(z, (u64)t * SCAN_PER_THREAD));
    __syncthreads();
    for (u32 s = SCAN_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) red[t] = ext_add(red[t], red[t + s]);
        __syncthreads();
    }
    if (t == 0) tot[blockIdx.x] = red[0];
}

// carry[s] = S[end of segment s] = sum_{s' > s} tot[s'] z^(SCAN_SEG (s' - s - 1)).  One CTA: every thread owns a run of
// `per` consecutive segments, the runs are combined with the same doubling scan as below.
__global__ void __launch_bounds__(SCAN_THREADS)
scan_carries_kernel(const Ext2* __restrict__ tot, u32 n_seg, Ext2 z, Ext2* __restrict__ carry) {
    __shared__ Ext2 V[SCAN_THREADS + 1];
    const u32 t = threadIdx.x;
    const u32 per = (n_seg + SCAN_THREADS - 1) / SCAN_THREADS;
    const u32 s0 = t * per, s1 = (s0 + per < n_seg) ? s0 + per : n_seg;
    const Ext2 step = ext_pow(z, SCAN_SEG);
    Ext2 h = ext_zero();
    for (u32 s = s1; s-- > s0 && s < n_seg;) h = ext_add(ext_mul(h, step), tot[s]);     // (empty when s0 >= n_seg)
    V[t] = h;
    if (t == 0) V[SCAN_THREADS] = ext_zero();
    __syncthreads();
    Ext2 zp = ext_pow(step, per);
    for (u32 d = 1; d < SCAN_THREADS; d <<= 1) {
        Ext2 v = V[t];
        if (t + d < SCAN_THREADS) v = ext_add(v, ext_mul(zp, V[t + d]));
        __syncthreads();
        V[t] = v;
        __syncthreads();
        zp = ext_mul(zp, zp);
    }
    Ext2 c = V[t + 1];                       // S at the end of this thread's run
    for (u32 s = s1; s-- > s0 && s < n_seg;) {
        carry[s] = c;
        c = ext_add(ext_mul(c, step), tot[s]);
    }
}

// PolynomialCoeffs::divide_by_linear fused with ReducingFactor::shift_poly and the `+=`:
//   S[i] = sum_{j >= i} c[j] z^(j - i)  (suffix Horner), quotient q[i] = S[i + 1] (i < n - 1), q[n - 1] = 0;
//   acc[i + off] = acc[i + off] * acc_scale + q[i],  off = mul_by_x (the 2022 plonky2 multiplies final_poly by X:
//   acc[0] stays 0 and the padding zero q[n - 1] falls off the end).
__global__ void __launch_bounds__(SCAN_THREADS)
divide_by_linear_kernel(const u64* __restrict__ in_a, const u64* __restrict__ in_b, u64 n, Ext2 z,
                        const Ext2* __restrict__ carry, Ext2 acc_scale, u32 mul_by_x,
                        u64* __restrict__ acc_a, u64* __restrict__ acc_b) {
    __shared__ Ext2 V[SCAN_THREADS + 1];
    const u32 t = threadIdx.x;
    const u64 base = (u64)blockIdx.x * SCAN_SEG + (u64)t * SCAN_PER_THREAD;
    Ext2 h[SCAN_PER_THREAD + 1];
    h[SCAN_PER_THREAD] = ext_zero();
#pragma unroll
    for (int e = SCAN_PER_THREAD - 1; e >= 0; e--) {
        Ext2 c = ext_zero();
        if (base + e < n) { c.a = gl::canon(in_a[base + e]); c.b = gl::canon(in_b[base + e]); }
        h[e] = ext_add(ext_mul(h[e + 1], z), c);
    }
    // V[t] = sum_{u >= t} h_u[0] z^(E (u - t)) over this CTA (Hillis-Steele, multiplier doubles each step)
    V[t] = h[0];
    if (t == 0) V[SCAN_THREADS] = ext_zero();
    __syncthreads();
    Ext2 zp = ext_pow(z, SCAN_PER_THREAD);
    for (u32 d = 1; d < SCAN_THREADS; d <<= 1) {
        Ext2 v = V[t];
        if (t + d < SCAN_THREADS) v = ext_add(v, ext_mul(zp, V[t + d]));
        __syncthreads();
        V[t] = v;
        __syncthreads();
        zp = ext_mul(zp, zp);
    }
    // W = S at the start of the next thread's chunk = V[t + 1] + z^(E (255 - t)) * carry
    Ext2 W = ext_add(V[t + 1], ext_mul(ext_pow(z, (u64)(SCAN_THREADS - 1 - t) * SCAN_PER_THREAD), carry[blockIdx.x]));
    Ext2 zk = ext_zero(); zk.a = 1;     // z^(E - e - 1), built from e = E - 1 downwards
#pragma unroll
    for (int e = SCAN_PER_THREAD - 1; e >= 0; e--) {
        u64 i = base + e;
        Ext2 q = ext_add(h[e + 1], ext_mul(zk, W));     // S[i + 1]
        zk = ext_mul(zk, z);
        u64 o = i + mul_by_x;
        if (i < n && o < n) {
            Ext2 a; a.a = acc_a[o]; a.b = acc_b[o];
            a = ext_add(ext_mul(a, acc_scale), q);
            acc_a[o] = a.a; acc_b[o] = a.b;
        }
    }
    if (mul_by_x && blockIdx.x == 0 && t == 0) {        // X * (...) has no constant term
        Ext2 a; a.a = acc_a[0]; a.b = acc_b[0];
        a = ext_mul(a, acc_scale);
        acc_a[0] = a.a; acc_b[0] = a.b;
    }
}

// planes [2][len] -> interleaved rows [len][2]  (plonky2 `flatten` of a chunk = 2 * arity consecutive words)
__global__ void interleave_kernel(const u64* __restrict__ pa, const u64* __restrict__ pb, u64 len, u64* __restrict__ out) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;   // output word
    if (g >= 2 * len) return;
    out[g] = (g & 1) ? pb[g >> 1] : pa[g >> 1];
}
__global__ void deinterleave_kernel(const u64* __restrict__ in, u64 len, u64* __restrict__ pa, u64* __restrict__ pb) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 2 * len) return;
    u64 v = gl::canon(in[g]);
    if (g & 1) pb[g >> 1] = v; else pa[g >> 1] = v;
}

// fri_committed_trees fold: out[j] = reduce_with_powers(c[j*arity .. (j+1)*arity), beta) = sum_i beta^i c[j*arity + i]
__global__ void fold_kernel(const u64* __restrict__ in_a, const u64* __restrict__ in_b, u64 out_len, u32 arity, Ext2 beta,
                            u64* __restrict__ out_a, u64* __restrict__ out_b) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= out_len) return;
    Ext2 acc = ext_zero();
    for (u32 i = arity; i-- > 0;) {
        Ext2 c; c.a = in_a[j * arity + i]; c.b = in_b[j * arity + i];
        acc = ext_add(ext_mul(acc, beta), c);
    }
    out_a[j] = acc.a;
    out_b[j] = acc.b;
}

// rows of a row-major table: out[q][c] = src[idx[q]][c]
__global__ void gather_rows_rm_kernel(const u64* __restrict__ src, u32 row_len, const u64* __restrict__ idx, u64 n_idx,
                                      u64* __restrict__ out) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_idx * row_len) return;
    u64 q = g / row_len, c = g % row_len;
    out[g] = src[idx[q] * row_len + c];
}

// fri_proof_of_work: smallest candidate w in [start, start + count) such that
//   permute(state with state[witness_pos] = w)[response_pos]  has >= min_lz leading zero bits (canonical u64).
// Covers both published forms: hash_no_pad(current_hash || w).elements[0]  (state = hash || 0.., pos 4, response 0)
// and the duplex form (state = sponge state overwritten by the input buffer, pos = buffer length, response 7).
__global__ void __launch_bounds__(256)
pow_grind_kernel(const u64* __restrict__ state, u32 witness_pos, u32 response_pos, u32 min_lz, u64 start, u64 count,
                 unsigned long long* __restrict__ best) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    u64 w = start + g;
    u64 s[poseidon::WIDTH];
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) s[i] = (u32)i == witness_pos ? w : state[i];
    poseidon::permute(s);
    u64 r = 0;
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) if ((u32)i == response_pos) r = gl::canon(s[i]);
    u32 lz = r ? (u32)__clzll((long long)r) : 64u;
    if (lz >= min_lz) atomicMin(best, (unsigned long long)w);
}
#endif

}  // namespace frik
