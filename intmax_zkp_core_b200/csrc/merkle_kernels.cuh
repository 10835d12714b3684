// Poseidon leaf hashing, Merkle level reduction and the small batched Hasher kernels.
//
// Replaces plonky2 `MerkleTree::new` / `fill_digests_buf` / `fill_subtree`
// (plonky2 @ f99ed9c, plonky2/src/hash/merkle_tree.rs), `hash_n_to_m_no_pad`, `compress`
// (plonky2/src/hash/hashing.rs) and `Hasher::hash_or_noop`; SURVEY.md rows A6-A8, A11.
// Reached from the reference through every prove()/build(), e.g.
// /root/reference/src/rollup/circuits/mod.rs:1247 and :605.
//
// One thread owns one sponge (12-word state in registers).  Leaves are read through generic
// (row_stride, col_stride) so the same kernel serves the column-major LDE the NTT writes (coalesced:
// consecutive threads = consecutive rows of one column) and row-major leaves given by a caller.
// Digests are written straight into plonky2's interleaved layout
//   subtree s at [s*L, (s+1)*L), L = 2*(leaves_per_subtree - 1);
//   node (layer i, index m): 2*(((m>>1) << (i+1)) + (1<<i) - 1) + (m&1);   layer 0 = leaf digests
// and subtree roots into the cap.
#pragma once
#include "poseidon.cuh"

#ifndef B200ZKP_HASH_MINBLOCKS
#define B200ZKP_HASH_MINBLOCKS 2
#endif
#ifndef B200ZKP_HASH_THREADS
#define B200ZKP_HASH_THREADS 512
#endif
#ifndef B200ZKP_HASH_PREFETCH
#define B200ZKP_HASH_PREFETCH 0
#endif

namespace merkle {

using gl::u32;
using gl::u64;

struct TreeShape {
    u32 sub_log;      // log2(leaves per cap subtree) = log2(N) - cap_height
    u64 sub_digests;  // 2 * (2^sub_log - 1)
};

GL_FN u64 node_slot(const TreeShape& t, u64 subtree, u32 layer, u64 m) {
    return subtree * t.sub_digests + 2 * (((m >> 1) << (layer + 1)) + ((u64)1 << layer) - 1) + (m & 1);
}

// hash_or_noop (noop_short = 1) or hash_no_pad (0) of one leaf; element c of the leaf at p[c*col_stride].
GL_FN void sponge_leaf(const u64* __restrict__ p, u64 col_stride, u32 leaf_len, u32 noop_short,
                       u64 (&s)[poseidon::WIDTH]) {
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) s[i] = 0;
    if (leaf_len <= 4 && noop_short) {
        for (u32 i = 0; i < leaf_len; i++) {
            u64 v = gl::canon(p[i * col_stride]);
#pragma unroll
            for (int q = 0; q < 4; q++) if (q == (int)i) s[q] = v;
        }
        return;
    }
#if !B200ZKP_HASH_PREFETCH
    for (u32 c = 0; c < leaf_len; c += poseidon::RATE) {
#pragma unroll
        for (int i = 0; i < poseidon::RATE; i++)
            if (c + i < leaf_len) s[i] = p[(u64)(c + i) * col_stride];
        poseidon::permute(s);
    }
#else
    u64 nx[poseidon::RATE];
#pragma unroll
    for (int i = 0; i < poseidon::RATE; i++) nx[i] = (u32)i < leaf_len ? p[(u64)i * col_stride] : 0;
    for (u32 c = 0; c < leaf_len; c += poseidon::RATE) {
        u32 rem = leaf_len - c;
        // overwrite mode: a short last chunk leaves the remaining rate words untouched
#pragma unroll
        for (int i = 0; i < poseidon::RATE; i++) if ((u32)i < rem) s[i] = nx[i];
        // prefetch the next chunk before the ~20k-instruction permutation
        u32 c2 = c + poseidon::RATE;
#pragma unroll
        for (int i = 0; i < poseidon::RATE; i++)
            if (c2 + i < leaf_len) nx[i] = p[(u64)(c2 + i) * col_stride];
        poseidon::permute(s);
    }
#endif
}

#ifndef B200ZKP_HOST_EMU
__device__ __forceinline__ void store_digest(u64* dst, const u64 (&s)[poseidon::WIDTH]) {
    ulonglong2* d = reinterpret_cast<ulonglong2*>(dst);   // digests are 32-byte aligned
    d[0] = make_ulonglong2(s[0], s[1]);
    d[1] = make_ulonglong2(s[2], s[3]);
}

// hash_or_noop over one leaf per thread.  leaf element (row, c) = leaves[row*row_stride + c*col_stride].
__global__ void __launch_bounds__(B200ZKP_HASH_THREADS, B200ZKP_HASH_MINBLOCKS)
leaf_hash_kernel(const u64* __restrict__ leaves, u64 row_stride, u64 col_stride, u32 leaf_len,
                 u64 row0, u64 n_rows, TreeShape shape, u64* __restrict__ digests, u64* __restrict__ cap,
                 u32 noop_short /* 1: hash_or_noop, 0: hash_no_pad */) {
    // rows [row0, row0 + n_rows) of the leaf range: one launch per finished coset block when pipelined with the LDE
    u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_rows) return;
    u64 row = row0 + gid;
    u64 s[poseidon::WIDTH];
    sponge_leaf(leaves + row * row_stride, col_stride, leaf_len, noop_short, s);
    if (shape.sub_log == 0) {
        store_digest(cap + 4 * row, s);
    } else {
        u64 subtree = row >> shape.sub_log;
        u64 m = row & (((u64)1 << shape.sub_log) - 1);
        store_digest(digests + 4 * node_slot(shape, subtree, 0, m), s);
    }
}

// parents of layer `layer` (children) -> layer+1, or the cap when layer+1 == sub_log.
__global__ void __launch_bounds__(B200ZKP_HASH_THREADS, B200ZKP_HASH_MINBLOCKS)
merkle_level_kernel(u64* __restrict__ digests, u64* __restrict__ cap, TreeShape shape, u32 layer,
                    u64 n_parents) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_parents) return;
    u32 par_log = shape.sub_log - layer - 1;           // log2(parents per subtree)
    u64 subtree = g >> par_log;
    u64 m = g & (((u64)1 << par_log) - 1);
    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(digests + 4 * node_slot(shape, subtree, layer, 2 * m));
    u64 s[poseidon::WIDTH];
    ulonglong2 a = src[0], b = src[1], c = src[2], d = src[3];
    s[0] = a.x; s[1] = a.y; s[2] = b.x; s[3] = b.y; s[4] = c.x; s[5] = c.y; s[6] = d.x; s[7] = d.y;
    s[8] = s[9] = s[10] = s[11] = 0;
    poseidon::permute(s);
    if (par_log == 0) store_digest(cap + 4 * subtree, s);
    else store_digest(digests + 4 * node_slot(shape, subtree, layer + 1, m), s);
}

// ---------------------------------------------------------------- stateless Hasher helpers (tests, N4)
// out[i] = permute(in[i]), states row-major 12 words each
__global__ void __launch_bounds__(128)
permute_kernel(const u64* __restrict__ in, u64* __restrict__ out, u64 count) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    u64 s[poseidon::WIDTH];
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) s[i] = in[g * poseidon::WIDTH + i];
    poseidon::permute(s);
#pragma unroll
    for (int i = 0; i < poseidon::WIDTH; i++) out[g * poseidon::WIDTH + i] = s[i];
}

// out[i] = two_to_one(l[i], r[i])
__global__ void __launch_bounds__(128)
two_to_one_kernel(const u64* __restrict__ l, const u64* __restrict__ r, u64* __restrict__ out, u64 count) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    u64 s[poseidon::WIDTH];
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] = l[4 * g + i]; s[4 + i] = r[4 * g + i]; s[8 + i] = 0; }
    poseidon::permute(s);
    store_digest(out + 4 * g, s);
}

// ---------------------------------------------------------------- accessors (A11)
// rows[q][c] = lde[c*col_stride + idx[q]]  (leaf rows of a column-major LDE)
__global__ void gather_rows_kernel(const u64* __restrict__ lde, u64 col_stride, u32 row_len,
                                   const u64* __restrict__ idx, u64 n_idx, u64* __restrict__ rows) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_idx * row_len) return;
    u64 q = g / row_len;
    u32 c = (u32)(g % row_len);
    rows[g] = lde[(u64)c * col_stride + idx[q]];
}

// MerkleTree::prove for a batch of leaf indices: siblings[q][i] = sibling digest at layer i
__global__ void gather_siblings_kernel(const u64* __restrict__ digests, TreeShape shape,
                                       const u64* __restrict__ idx, u64 n_idx, u64* __restrict__ sib) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_idx * shape.sub_log) return;
    u64 q = g / shape.sub_log;
    u32 layer = (u32)(g % shape.sub_log);
    u64 leaf = idx[q];
    u64 subtree = leaf >> shape.sub_log;
    u64 m = (leaf & (((u64)1 << shape.sub_log) - 1)) >> layer;   // node on the path at this layer
    const u64* src = digests + 4 * node_slot(shape, subtree, layer, m ^ 1);
#pragma unroll
    for (int i = 0; i < 4; i++) sib[4 * g + i] = src[i];
}

// column-major [cols][col_stride] -> row-major [rows][cols] for rows in [row0, row0 + n_rows)
__global__ void transpose_to_rows_kernel(const u64* __restrict__ cm, u64 col_stride, u32 cols,
                                         u64 row0, u64 n_rows, u64* __restrict__ rm) {
    __shared__ u64 tile[32][33];
    u64 rbase = (u64)blockIdx.x * 32;
    u32 cbase = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        u32 c = cbase + j;
        u64 r = rbase + threadIdx.x;
        if (c < cols && r < n_rows) tile[j][threadIdx.x] = cm[(u64)c * col_stride + row0 + r];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        u64 r = rbase + j;
        u32 c = cbase + threadIdx.x;
        if (c < cols && r < n_rows) rm[r * cols + c] = tile[threadIdx.x][j];
    }
}

// row-major [rows][cols] -> column-major [cols][col_stride]
__global__ void transpose_to_cols_kernel(const u64* __restrict__ rm, u32 cols, u64 n_rows,
                                         u64* __restrict__ cm, u64 col_stride) {
    __shared__ u64 tile[32][33];
    u64 rbase = (u64)blockIdx.x * 32;
    u32 cbase = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        u64 r = rbase + j;
        u32 c = cbase + threadIdx.x;
        if (c < cols && r < n_rows) tile[j][threadIdx.x] = rm[r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        u32 c = cbase + j;
        u64 r = rbase + threadIdx.x;
        if (c < cols && r < n_rows) cm[(u64)c * col_stride + r] = tile[threadIdx.x][j];
    }
}

#endif  // !B200ZKP_HOST_EMU

}  // namespace merkle
