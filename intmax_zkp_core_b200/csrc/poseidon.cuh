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
GL_FN u64 add_const(u64 a, u64 c) { return gl::add_nc(a, c); }

// value = L + H * 2^32 with L, H < 2^44  ->  arbitrary u64 congruent mod p
//   = l0 + (l1 + h0) B + h1 B^2,  B^2 == B - 1   (6 SASS instructions)
GL_FN u64 fold_lh(u64 L, u64 H) {
#ifdef B200ZKP_HOST_EMU
    u64 t = H << 32;
    u64 r = L + t;
    u64 top = (H >> 32) + (r < t ? 1u : 0u);      // multiples of 2^64, < 2^13
    u64 e = (top << 32) - top;                    // top * EPS
    u64 r2 = r + e;
    return (r2 < e) ? r2 + gl::EPS : r2;
#else
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, h0, h1, m0, e, m;\n\t"
        ".reg .u64 t, u;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "mov.b64 {h0, h1}, %2;\n\t"
        "add.cc.u32 m0, l1, h0;\n\t"
        "addc.u32 e, h1, 0;\n\t"               // multiples of 2^64
        "mov.b64 t, {l0, m0};\n\t"
        "mul.wide.u32 u, e, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, t, u;\n\t"
        "addc.u32 m, 0, 0;\n\t"               // m = carry (0/1); NB: subc after add.cc has the wrong polarity
        "mov.b64 {l0, l1}, t;\n\t"
        "sub.cc.u32 l0, l0, m;\n\t"           // + m * EPS == + m * 2^32 - m (cannot carry twice)
        "subc.u32 l1, l1, 0;\n\t"
        "add.u32 l1, l1, m;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(L), "l"(H));
    return r;
#endif
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
            // diag(8) folded into the circulant entry of word 0 for row 0: 17 + 8
            const u32 c = (r == 0 && i == 0) ? C[0] + 8u : C[i];
            L = gl::mad_wide(lo[(i + r) % WIDTH], c, L);
            H = gl::mad_wide(hi[(i + r) % WIDTH], c, H);
        }
        s[r] = fold_lh(L, H);
    }
}

// Column accumulator for dot products of 64-bit words: sum of 32x32 partial products of one weight,
// as a 64-bit sum plus a carry count (one IMAD.WIDE.U32 with carry-out + one add-with-carry per term).
struct Col {
    u64 acc;
    u32 cnt;
    GL_MFN void init() { acc = 0; cnt = 0; }
    GL_MFN void mac(u32 a, u32 b) {
#ifdef B200ZKP_HOST_EMU
        u64 p = (u64)a * b;
        acc += p;
        cnt += acc < p ? 1u : 0u;
#else
        asm("{\n\t"
            ".reg .u64 u;\n\t"
            "mul.wide.u32 u, %2, %3;\n\t"
            "add.cc.u64 %0, %0, u;\n\t"
            "addc.u32 %1, %1, 0;\n\t"
            "}" : "+l"(acc), "+r"(cnt) : "r"(a), "r"(b));
#endif
    }
};

// D = C0 + C1*B + C2*B^2 (B = 2^32) with C_i = cnt_i*B^2 + acc_i  ->  u64 congruent mod p
GL_FN u64 reduce_cols(const Col& c0, const Col& c1, const Col& c2) {
    // limbs: c00 + (c01 + c10) B + (c02 + c11 + c20) B^2 + (c12 + c21) B^3 + c22 B^4
    u32 c00 = (u32)c0.acc, c01 = (u32)(c0.acc >> 32), c02 = c0.cnt;
    u32 c10 = (u32)c1.acc, c11 = (u32)(c1.acc >> 32), c12 = c1.cnt;
    u32 c20 = (u32)c2.acc, c21 = (u32)(c2.acc >> 32), c22 = c2.cnt;
    // carry-propagate into a 160-bit integer (x4 : x3 : x2 : x1 : x0), x4 < 2^6
    u64 t1 = (u64)c01 + c10;
    u64 t2 = (u64)c02 + c11 + c20 + (t1 >> 32);
    u64 t3 = (u64)c12 + c21 + (t2 >> 32);
    u32 x4 = c22 + (u32)(t3 >> 32);
    u64 lo = ((u64)(u32)t1 << 32) | c00;
    u64 hi = ((u64)(u32)t3 << 32) | (u32)t2;
    // B^4 == -B: subtract x4 * 2^32 (canonical, < 2^38) after reducing the low 128 bits
    u64 r = gl::canon(gl::reduce128(lo, hi));
    return gl::sub(r, (u64)x4 << 32);
}

// s + w * x  (all arbitrary u64) -> arbitrary u64
GL_FN u64 mul_add_nc(u64 w, u64 x, u64 s) {
    unsigned __int128 p = (unsigned __int128)w * x + s;   // <= (2^64-1)^2 + 2^64 - 1 < 2^128
    return gl::reduce128((u64)p, (u64)(p >> 64));
}

GL_FN u64 dot12(const u64* __restrict__ w, u64 w0const, const u64 (&s)[WIDTH], u64 s0) {
    // w0const * s0 + sum_j w[j] * s[1+j]
    Col c0, c1, c2;
    c0.init(); c1.init(); c2.init();
    {
        u32 a0 = (u32)s0, a1 = (u32)(s0 >> 32), b0 = (u32)w0const;   // w0const < 2^32
        c0.mac(a0, b0); c1.mac(a1, b0);
    }
#pragma unroll
    for (int j = 0; j < WIDTH - 1; j++) {
        u64 wj = w[j];
        u32 b0 = (u32)wj, b1 = (u32)(wj >> 32);
        u32 a0 = (u32)s[1 + j], a1 = (u32)(s[1 + j] >> 32);
        c0.mac(a0, b0); c1.mac(a0, b1); c1.mac(a1, b0); c2.mac(a1, b1);
    }
    return reduce_cols(c0, c1, c2);
}

GL_FN void partial_rounds_fast(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
    // dense 11x11 on words 1..11 (once)
    {
        u64 t[WIDTH - 1];
#pragma unroll
        for (int r = 0; r < WIDTH - 1; r++) t[r] = dot12(&FAST_INIT[r * (WIDTH - 1)], 0, s, 0);
#pragma unroll
        for (int q = 0; q < WIDTH - 1; q++) s[1 + q] = t[q];
    }
#pragma unroll 1
    for (int i = 0; i < 22; i++) {
        u64 s0 = add_const(sbox(s[0]), FAST_POST[i]);
        u64 d = dot12(&FAST_VHAT[i * (WIDTH - 1)], 25, s, s0);
#pragma unroll
        for (int j = 0; j < WIDTH - 1; j++) s[1 + j] = mul_add_nc(FAST_WHAT[i * (WIDTH - 1) + j], s0, s[1 + j]);
        s[0] = d;
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
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], RC_FULL_PAD[4 * WIDTH + i]);
#else
    partial_rounds_pushed(s);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], PUSH_TAIL[i]);
#endif
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        // the last round adds nothing: rows 8.. of RC_FULL_PAD are zero, so one MDS body serves all rounds
        mds_layer<true>(s, &RC_FULL_PAD[(r + 1) * WIDTH]);
    }
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::canon(s[i]);
}

}  // namespace poseidon
