// Poseidon-Goldilocks permutation, one 12-word state per thread, state kept in registers.
//
// Replaces plonky2 `PoseidonPermutation::permute` / `Poseidon::poseidon` for GoldilocksField
// (plonky2 @ f99ed9c, plonky2/src/hash/poseidon.rs + poseidon_goldilocks.rs; reached from the
// reference through PoseidonHash::two_to_one / hash_pad, e.g.
// /root/reference/src/sparse_merkle_tree/goldilocks_poseidon/mod.rs:161-183, and through every
// MerkleTree::new inside prove(), /root/reference/src/transaction/circuits/mod.rs:453).
//
// Layout of the work: 4 full rounds, 22 partial rounds, 4 full rounds, S-box x^7,
// MDS = circulant(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...).
//   * full-round MDS: each word is split into 32-bit halves; 12x12 small-constant MACs per half are
//     IMAD.WIDE.U32 accumulations (sums < 2^42), recombined with one Solinas fold.  The NEXT round's
//     constants are folded into the accumulators, so round-constant addition is free.
//   * partial rounds: either the "fast" sparse form (22 64-bit MACs per round, tables re-derived in
//     tools/poseidon_derive.py) or the "pushed" form (scalar constant on word 0 + small-constant MDS).
// Words between rounds are arbitrary u64 (not canonical); `permute` canonicalises its output.
#pragma once
#include "goldilocks.cuh"
#include "poseidon_tables.cuh"

namespace poseidon {

using gl::u32;
using gl::u64;

static constexpr int WIDTH = 12;
static constexpr int RATE = 8;

#ifndef B200ZKP_FAST_PARTIAL
#define B200ZKP_FAST_PARTIAL 1
#endif

GL_FN u64 sbox(u64 x) {
    u64 x2 = gl::mul_nc(x, x);
    u64 x4 = gl::mul_nc(x2, x2);
    u64 x3 = gl::mul_nc(x2, x);
    return gl::mul_nc(x3, x4);
}

// a arbitrary u64, c canonical constant -> arbitrary u64 congruent to a + c
GL_FN u64 add_const(u64 a, u64 c) {
    u64 s = a + c;
    return (s < a) ? s + gl::EPS : s;   // single wrap: cannot wrap twice because c < p
}

// value = L + H * 2^32 with L, H < 2^44  ->  arbitrary u64 congruent mod p
GL_FN u64 fold_lh(u64 L, u64 H) {
    u64 t = H << 32;
    u64 r = L + t;
    u64 top = (H >> 32) + (r < t ? 1u : 0u);      // multiples of 2^64, < 2^13
    u64 e = (top << 32) - top;                    // top * EPS
    u64 r2 = r + e;
    return (r2 < e) ? r2 + gl::EPS : r2;
}

// s <- MDS * s (+ addc, the constants of the next round, if addc != nullptr)
template <bool kAddConst>
GL_FN void mds_layer(u64 (&s)[WIDTH], const unsigned long long* addc) {
    constexpr u32 C[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[WIDTH], hi[WIDTH];
#pragma unroll
    for (int i = 0; i < WIDTH; i++) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) {
        u64 L = 0, H = 0;
        if (kAddConst) { u64 c = addc[r]; L = (u32)c; H = c >> 32; }
#pragma unroll
        for (int i = 0; i < WIDTH; i++) {
            L += (u64)lo[(i + r) % WIDTH] * C[i];
            H += (u64)hi[(i + r) % WIDTH] * C[i];
        }
        if (r == 0) { L += (u64)lo[0] * 8u; H += (u64)hi[0] * 8u; }
        s[r] = fold_lh(L, H);
    }
}

// 128-bit accumulator with overflow count, for dot products of 64-bit words (<= 2^8 terms)
struct Acc160 {
    u64 lo, hi;
    u32 top;
    GL_MFN void init() { lo = 0; hi = 0; top = 0; }
    GL_MFN void mac(u64 a, u64 b) {
        unsigned __int128 p = (unsigned __int128)a * b;
        u64 pl = (u64)p, ph = (u64)(p >> 64);
        lo += pl;
        u64 c = lo < pl ? 1u : 0u;
        u64 h2 = hi + ph;
        u32 c2 = h2 < ph ? 1u : 0u;
        u64 h3 = h2 + c;
        c2 += h3 < h2 ? 1u : 0u;
        hi = h3;
        top += c2;
    }
    // 2^128 == -2^32 (mod p)
    GL_MFN u64 reduce() const {
        u64 r = gl::canon(gl::reduce128(lo, hi));
        u64 t = (u64)top << 32;                    // < 2^40, canonical
        return gl::sub(r, t);
    }
};

// s[j] + w * x  (s arbitrary u64) -> arbitrary u64
GL_FN u64 mul_add_nc(u64 w, u64 x, u64 s) {
    unsigned __int128 p = (unsigned __int128)w * x + s;   // <= (2^64-1)^2 + 2^64 - 1 < 2^128
    return gl::reduce128((u64)p, (u64)(p >> 64));
}

GL_FN void partial_rounds_fast(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
    // dense 11x11 on words 1..11 (once)
    {
        u64 t[WIDTH - 1];
#pragma unroll
        for (int r = 0; r < WIDTH - 1; r++) {
            Acc160 acc; acc.init();
#pragma unroll
            for (int c = 0; c < WIDTH - 1; c++) acc.mac(FAST_INIT[r * (WIDTH - 1) + c], s[1 + c]);
            t[r] = acc.reduce();
        }
#pragma unroll
        for (int q = 0; q < WIDTH - 1; q++) s[1 + q] = t[q];
    }
#pragma unroll 1
    for (int i = 0; i < 22; i++) {
        u64 s0 = add_const(sbox(s[0]), FAST_POST[i]);
        Acc160 acc; acc.init();
        acc.mac(s0, 25);
#pragma unroll
        for (int j = 0; j < WIDTH - 1; j++) acc.mac(FAST_VHAT[i * (WIDTH - 1) + j], s[1 + j]);
#pragma unroll
        for (int j = 0; j < WIDTH - 1; j++) s[1 + j] = mul_add_nc(FAST_WHAT[i * (WIDTH - 1) + j], s0, s[1 + j]);
        s[0] = acc.reduce();
    }
}

GL_FN void partial_rounds_pushed(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
#pragma unroll 1
    for (int i = 0; i < 22; i++) {
        s[0] = sbox(add_const(s[0], PUSH_SCAL[i]));
        mds_layer<false>(s, nullptr);
    }
}

// In-place permutation; input words arbitrary u64, output canonical.
GL_FN void permute(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], RC_FULL[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
#if B200ZKP_FAST_PARTIAL
        const unsigned long long* nxt = (r < 3) ? &RC_FULL[(r + 1) * WIDTH] : FAST_FIRST;
        mds_layer<true>(s, nxt);
#else
        if (r < 3) mds_layer<true>(s, &RC_FULL[(r + 1) * WIDTH]);
        else mds_layer<false>(s, nullptr);
#endif
    }
#if B200ZKP_FAST_PARTIAL
    partial_rounds_fast(s);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], RC_FULL[4 * WIDTH + i]);
#else
    partial_rounds_pushed(s);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], PUSH_TAIL[i]);
#endif
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        if (r < 7) mds_layer<true>(s, &RC_FULL[(r + 1) * WIDTH]);
        else mds_layer<false>(s, nullptr);
    }
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::canon(s[i]);
}

}  // namespace poseidon
