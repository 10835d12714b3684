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
#define B200ZKP_HASH_MINBLOCKS 3
#endif
#ifndef B200ZKP_HASH_THREADS
#define B200ZKP_HASH_THREADS 256
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
    for (u32 c = 0; c < leaf_len; c += poseidon::RATE) {
        // overwrite mode: a short last chunk leaves the remaining rate words untouched
#pragma unroll
        for (int i = 0; i < poseidon::RATE; i++)
            if (c + i < leaf_len) s[i] = p[(u64)(c + i) * col_stride];
        poseidon::permute(s);
    }
}

// columns [c_begin, c_end) of one leaf absorbed into the running state of its sponge (c_begin a multiple of the rate; a short
// last chunk leaves the remaining rate words as the previous permutation left them: overwrite mode).  hash_no_pad of a leaf
// = this over [0, leaf_len) in any number of pieces cut at multiples of the rate; the digest is words 0..3 at the end.
GL_FN void sponge_absorb(const u64* __restrict__ p, u64 col_stride, u32 c_begin, u32 c_end, u64 (&s)[poseidon::WIDTH]) {
    for (u32 c = c_begin; c < c_end; c += poseidon::RATE) {
#pragma unroll
        for (int i = 0; i < poseidon::RATE; i++)
            if (c + i < c_end) s[i] = p[(u64)(c + i) * col_stride];
        poseidon::permute(s);
    }
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

// Resumable form of leaf_hash_kernel for column-major leaves that arrive column group by column group (the partitioned
// commitment with host inputs, sharded.inl): columns [c_begin, c_end) of every row are absorbed; between launches the 12-word
// state of row r lives at state[w * state_stride + r] (word-major: coalesced).  c_begin == 0 starts from the zero state,
// c_end == leaf_len (> 4: hash_or_noop hashes) writes the digest instead of the state.
__global__ void __launch_bounds__(B200ZKP_HASH_THREADS, B200ZKP_HASH_MINBLOCKS)
leaf_absorb_kernel(const u64* __restrict__ leaves, u64 col_stride, u32 leaf_len, u32 c_begin, u32 c_end, u64 n_rows, TreeShape shape,
                   u64* __restrict__ state, u64 state_stride, u64* __restrict__ digests, u64* __restrict__ cap) {
    const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    u64 s[poseidon::WIDTH];
    if (c_begin == 0) {
#pragma unroll
        for (int i = 0; i < poseidon::WIDTH; i++) s[i] = 0;
    } else {
#pragma unroll
        for (int i = 0; i < poseidon::WIDTH; i++) s[i] = state[(u64)i * state_stride + row];
    }
    sponge_absorb(leaves + row, col_stride, c_begin, c_end, s);
    if (c_end < leaf_len) {
#pragma unroll
        for (int i = 0; i < poseidon::WIDTH; i++) state[(u64)i * state_stride + row] = s[i];
        return;
    }
    if (shape.sub_log == 0) {
        store_digest(cap + 4 * row, s);
    } else {
        const u64 subtree = row >> shape.sub_log;
        const u64 m = row & (((u64)1 << shape.sub_log) - 1);
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
    poseidon::permute_digest(s);
    if (par_log == 0) store_digest(cap + 4 * subtree, s);
    else store_digest(digests + 4 * node_slot(shape, subtree, layer + 1, m), s);
}

// ---------------------------------------------------------------- latency form of the permutation
// One state per 16-lane group, word i in lane i (lanes 12..15 idle), for launches too small to fill the GPU: the top
// levels of every tree, single compressions, the Fiat-Shamir transcript.  A thread of the throughput form runs ~22 k
// dependent-ish instructions per permutation (~25 us alone on an SM); here a lane runs one S-box and one MDS row per round
// (~150 instructions, the 11 foreign words arrive by warp shuffle), ~4x less latency for 3.6x more total work — a good
// trade only while the machine is otherwise idle (launch_tree_levels switches at 2048 nodes).
// Same arithmetic as poseidon::permute (pushed-constant schedule ROUND_ADD), exact mod p, canonical output.
// `ra` = ROUND_ADD in global memory (a lane-indexed read of the __constant__ copy would serialise).
__device__ __forceinline__ u64 coop_permute(u64 s, u32 li, u32 grp_base, const u64* __restrict__ ra) {
    constexpr u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    s = poseidon::add_const(s, __ldg(ra + li));
    u64 nxt = __ldg(ra + 12 + li);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        if (full || li == 0) s = poseidon::sbox(s);
        // out[l] = sum_i in[(l + i) mod 12] * CIRC[i] + (l == 0) * 8 in[0], accumulated on 32-bit halves (each sum < 2^43)
        u64 acc_lo = 0, acc_hi = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            u32 j = li + i;
            j = j >= 12 ? j - 12 : j;
            u64 v = __shfl_sync(0xffffffffu, s, (int)(grp_base + j));
            acc_lo += (u64)(u32)v * CIRC[i];
            acc_hi += (v >> 32) * CIRC[i];
        }
        if (li == 0) { acc_lo += (u64)(u32)s * 8u; acc_hi += (s >> 32) * 8u; }
        u64 lo = acc_lo + (acc_hi << 32);
        u64 hi = (acc_hi >> 32) + (lo < acc_lo ? 1u : 0u);
        s = gl::reduce128(lo, hi);
        if (r + 1 < 30) {
            const bool next_full = (r + 1 < 4) || (r + 1 >= 26);
            if (next_full || li == 0) s = poseidon::add_const(s, nxt);
            if (r + 2 < 30) nxt = __ldg(ra + (r + 2) * 12 + li);
        }
    }
    return gl::canon(s);
}

// parents of layer `layer`, one node per 16-lane group (see coop_permute); blockDim.x a multiple of 32
__global__ void __launch_bounds__(128)
merkle_level_coop_kernel(u64* __restrict__ digests, u64* __restrict__ cap, TreeShape shape, u32 layer, u64 n_parents,
                         const u64* __restrict__ ra) {
    const u32 lane = threadIdx.x & 31, l = lane & 15, grp_base = lane & 16;
    const u32 li = l < 12 ? l : 0;
    u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool live = g < n_parents;
    if (!live) g = 0;                                   // idle groups recompute node 0 and store nothing (shuffles stay full-warp)
    u32 par_log = shape.sub_log - layer - 1;
    u64 subtree = g >> par_log;
    u64 m = g & (((u64)1 << par_log) - 1);
    const u64* src = digests + 4 * node_slot(shape, subtree, layer, 2 * m);   // left || right child: 8 consecutive words
    u64 s = (l < 8) ? src[l] : 0;
    s = coop_permute(s, li, grp_base, ra);
    if (live && l < 4) {
        u64* dst = (par_log == 0) ? cap + 4 * subtree : digests + 4 * node_slot(shape, subtree, layer + 1, m);
        dst[l] = s;
    }
}

// The top of every cap subtree in ONE launch: CTA s reduces subtree s from layer `layer0` (at most TOP_PARENTS parents per
// subtree) up to its cap entry, one node per 16-lane group, __syncthreads() between layers (a subtree never reads another
// subtree's digests, so no grid-wide step is needed).  Replaces the chain of up to 8 dependent latency-form launches that
// ended every tree (profiles/r1c_bench_launch_list.csv); warps without a live node skip the round so that the few live
// warps of the last layers own the issue slots.  `digests` is read and written by different threads of the CTA: no
// __restrict__ / read-only path on it.
static constexpr u32 TOP_PARENTS_LOG = 7;      // 128 parents per subtree = 2 rounds of the 64 groups of a 1024-thread CTA
__global__ void __launch_bounds__(1024)
merkle_top_kernel(u64* digests, u64* __restrict__ cap, TreeShape shape, u32 layer0, const u64* __restrict__ ra) {
    const u32 lane = threadIdx.x & 31, l = lane & 15, grp_base = lane & 16;
    const u32 li = l < 12 ? l : 0;
    const u32 grp = threadIdx.x >> 4, n_grp = blockDim.x >> 4;
    const u32 warp_grp0 = (threadIdx.x >> 5) << 1;
    const u64 subtree = blockIdx.x;
    for (u32 layer = layer0; layer < shape.sub_log; layer++) {
        const u32 par_log = shape.sub_log - layer - 1;
        const u32 n_par = 1u << par_log;
        for (u32 m0 = 0; m0 < n_par; m0 += n_grp) {
            if (m0 + warp_grp0 >= n_par) continue;                   // warp-uniform: neither group of this warp has a node
            u32 m = m0 + grp;
            const bool live = m < n_par;
            if (!live) m = 0;
            const u64* src = digests + 4 * node_slot(shape, subtree, layer, 2 * (u64)m);
            u64 s = (l < 8) ? src[l] : 0;
            s = coop_permute(s, li, grp_base, ra);
            if (live && l < 4) {
                u64* dst = (par_log == 0) ? cap + 4 * subtree : digests + 4 * node_slot(shape, subtree, layer + 1, m);
                dst[l] = s;
            }
        }
        __syncthreads();
    }
}

// hash_or_noop / hash_no_pad of one leaf per 16-lane group (small trees: FRI layers, tiny circuits); same arguments as
// leaf_hash_kernel
__global__ void __launch_bounds__(128)
leaf_hash_coop_kernel(const u64* __restrict__ leaves, u64 row_stride, u64 col_stride, u32 leaf_len, u64 row0, u64 n_rows,
                      TreeShape shape, u64* __restrict__ digests, u64* __restrict__ cap, u32 noop_short,
                      const u64* __restrict__ ra) {
    const u32 lane = threadIdx.x & 31, l = lane & 15, grp_base = lane & 16;
    const u32 li = l < 12 ? l : 0;
    u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool live = g < n_rows;
    if (!live) g = 0;
    const u64 row = row0 + g;
    const u64* p = leaves + row * row_stride;
    u64 s = 0;
    if (leaf_len <= 4 && noop_short) {
        if (l < leaf_len) s = gl::canon(p[(u64)l * col_stride]);
    } else {
        for (u32 c = 0; c < leaf_len; c += poseidon::RATE) {        // overwrite mode: a short last chunk keeps the old rate words
            if (l < poseidon::RATE && c + l < leaf_len) s = p[(u64)(c + l) * col_stride];
            s = coop_permute(s, li, grp_base, ra);
        }
    }
    if (live && l < 4) {
        if (shape.sub_log == 0) cap[4 * row + l] = s;
        else {
            u64 subtree = row >> shape.sub_log;
            u64 m = row & (((u64)1 << shape.sub_log) - 1);
            digests[4 * node_slot(shape, subtree, 0, m) + l] = s;
        }
    }
}

// out[q] = permute(in[q]) (mode 0, 12 words in and out) or two_to_one(in[q][0..8]) (mode 1, 8 words in, 4 out), one state
// per 16-lane group
__global__ void __launch_bounds__(128)
permute_coop_kernel(const u64* __restrict__ in, const u64* __restrict__ in2, u64* __restrict__ out, u64 count, u32 mode,
                    const u64* __restrict__ ra) {
    const u32 lane = threadIdx.x & 31, l = lane & 15, grp_base = lane & 16;
    const u32 li = l < 12 ? l : 0;
    u64 g = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool live = g < count;
    if (!live) g = 0;
    u64 s = 0;
    if (mode == 0) { if (l < 12) s = in[g * 12 + l]; }
    else if (l < 4) s = in[g * 4 + l];
    else if (l < 8) s = in2[g * 4 + (l - 4)];
    s = coop_permute(s, li, grp_base, ra);
    if (live) {
        if (mode == 0) { if (l < 12) out[g * 12 + l] = s; }
        else if (l < 4) out[g * 4 + l] = s;
    }
}

// The Fiat-Shamir transcript's sponge (plonky2 iop/challenger.rs `duplexing`), a whole chain of steps in one launch of one warp:
// in = state[12] || inputs[n_inputs]: every chunk of up to 8 inputs overwrites the head of the state and is followed by a
// permutation; then n_squeeze further permutations, each publishing the 8 rate words.  out = final state[12] || squeezed.
__global__ void __launch_bounds__(32)
duplex_chain_kernel(const u64* __restrict__ in, u64 n_inputs, u32 n_squeeze, u64* __restrict__ out, const u64* __restrict__ ra) {
    const u32 lane = threadIdx.x & 31, l = lane & 15, grp_base = lane & 16;
    const u32 li = l < 12 ? l : 0;
    u64 s = l < 12 ? in[l] : 0;
    for (u64 c = 0; c < n_inputs; c += poseidon::RATE) {
        if (l < poseidon::RATE && c + l < n_inputs) s = in[12 + c + l];
        s = coop_permute(s, li, grp_base, ra);
    }
    for (u32 q = 0; q < n_squeeze; q++) {
        s = coop_permute(s, li, grp_base, ra);
        if (lane < poseidon::RATE) out[12 + q * poseidon::RATE + lane] = s;
    }
    if (lane < 12) out[lane] = s;
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
    poseidon::permute_digest(s);
    store_digest(out + 4 * g, s);
}

// ---------------------------------------------------------------- accessors (A11)
// rows[q][c] = lde[c*col_stride + idx[q]]  (leaf rows of a column-major LDE)
// idx_mask = number of leaves - 1: device-side indices cannot be validated by the host, so they are reduced instead
__global__ void gather_rows_kernel(const u64* __restrict__ lde, u64 col_stride, u32 row_len,
                                   const u64* __restrict__ idx, u64 n_idx, u64* __restrict__ rows, u64 idx_mask) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_idx * row_len) return;
    u64 q = g / row_len;
    u32 c = (u32)(g % row_len);
    rows[g] = lde[(u64)c * col_stride + (idx[q] & idx_mask)];
}

// MerkleTree::prove for a batch of leaf indices: siblings[q][i] = sibling digest at layer i
__global__ void gather_siblings_kernel(const u64* __restrict__ digests, TreeShape shape,
                                       const u64* __restrict__ idx, u64 n_idx, u64* __restrict__ sib, u64 idx_mask) {
    u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_idx * shape.sub_log) return;
    u64 q = g / shape.sub_log;
    u32 layer = (u32)(g % shape.sub_log);
    u64 leaf = idx[q] & idx_mask;
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
