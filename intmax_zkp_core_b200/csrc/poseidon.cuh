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
#define B200ZKP_FAST_PARTIAL 0
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
GL_FN void mds_layer_halves(u64 (&s)[WIDTH], const unsigned long long* addc) {
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

// value = O0 + O1 * 2^22 + O2 * 2^43 + rc  with O_i < 2^31, rc canonical  ->  arbitrary u64 congruent mod p
GL_FN u64 combine3(u32 O0, u32 O1, u32 O2, u64 rc) {
#ifdef B200ZKP_HOST_EMU
    unsigned __int128 v = (unsigned __int128)O0 + ((unsigned __int128)O1 << 22) + ((unsigned __int128)O2 << 43) + rc;
    u64 lo = (u64)v;
    u64 top = (u64)(v >> 64);                     // < 2^12
    u64 e = (top << 32) - top;                    // top * EPS
    u64 r = lo + e;
    return (r < e) ? r + gl::EPS : r;
#else
    u64 r;
    asm("{\n\t"
        ".reg .u32 a, b, c, d, w0, w1, top, rc0, rc1, m;\n\t"
        ".reg .u64 t, u;\n\t"
        "mov.b64 {rc0, rc1}, %4;\n\t"
        "shl.b32 a, %2, 22;\n\t"
        "shr.u32 b, %2, 10;\n\t"
        "shl.b32 c, %3, 11;\n\t"
        "shr.u32 d, %3, 21;\n\t"
        "add.cc.u32 w0, %1, a;\n\t"
        "addc.cc.u32 w1, b, c;\n\t"
        "addc.u32 top, d, 0;\n\t"
        "add.cc.u32 w0, w0, rc0;\n\t"
        "addc.cc.u32 w1, w1, rc1;\n\t"
        "addc.u32 top, top, 0;\n\t"             // multiples of 2^64 == EPS
        "mov.b64 t, {w0, w1};\n\t"
        "mul.wide.u32 u, top, 0xFFFFFFFF;\n\t"
        "add.cc.u64 t, t, u;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "mov.b64 {w0, w1}, t;\n\t"
        "sub.cc.u32 w0, w0, m;\n\t"             // + m * EPS
        "subc.u32 w1, w1, 0;\n\t"
        "add.u32 w1, w1, m;\n\t"
        "mov.b64 %0, {w0, w1};\n\t"
        "}" : "=l"(r) : "r"(O0), "r"(O1), "r"(O2), "l"(rc));
    return r;
#endif
}

// MDS on three limbs (22 + 21 + 21 bits) so every product and 12-term sum fits 32 bits: the multiplies are
// 32-bit IMADs (2 pipe cycles) instead of IMAD.WIDE (4), and the 12x12 circulant is split by
// x^12 - 1 = (x^6 - 1)(x^6 + 1) into a cyclic and a negacyclic 6x6 product whose halved constants are
//   P = (15, 24, 18, 17, 40, 14)   and   Q = (2, -4, 16, 1, -1, -1)      (all +-powers of two in Q)
// out[k] = A[k] + B[k], out[k+6] = A[k] - B[k]; arithmetic is mod 2^32 (exact: true sums < 2^31).
template <bool kAddConst>
GL_FN void mds_layer(u64 (&s)[WIDTH], const unsigned long long* addc) {
    constexpr u32 Pc[6] = {15, 24, 18, 17, 40, 14};
    constexpr int Qc[6] = {2, -4, 16, 1, -1, -1};
    u32 o[3][WIDTH];
#pragma unroll
    for (int L = 0; L < 3; L++) {
        u32 l[WIDTH];
#pragma unroll
        for (int i = 0; i < WIDTH; i++)
            l[i] = (L == 0) ? ((u32)s[i] & 0x3FFFFFu) : (L == 1) ? ((u32)(s[i] >> 22) & 0x1FFFFFu) : (u32)(s[i] >> 43);
        u32 sp[6], sm[6];
#pragma unroll
        for (int i = 0; i < 6; i++) { sp[i] = l[i] + l[i + 6]; sm[i] = l[i] - l[i + 6]; }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            u32 A = 0, B = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                A += sp[i] * Pc[(k - i + 6) % 6];
                const int q = (k >= i) ? Qc[k - i] : -Qc[k - i + 6];
                B += sm[i] * (u32)q;
            }
            o[L][k] = A + B;
            o[L][k + 6] = A - B;
        }
        o[L][0] += 8u * l[0];     // diag(8, 0, ..., 0)
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) s[r] = combine3(o[0][r], o[1][r], o[2][r], kAddConst ? (u64)addc[r] : 0ull);
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
#if B200ZKP_FAST_PARTIAL
GL_FN void permute(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], RC_FULL[i]);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        const unsigned long long* nxt = (r < 3) ? &RC_FULL[(r + 1) * WIDTH] : FAST_FIRST;
        mds_layer<true>(s, nxt);
    }
    partial_rounds_fast(s);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], RC_FULL_PAD[4 * WIDTH + i]);
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
#else
// One loop over all 30 rounds with a single copy of the S-box row and of the MDS body: the whole permutation is
// ~22 KB of SASS and stays resident in the instruction cache (the two-loop form was 59 KB and ncu showed
// "no instruction" as the top stall with a 67 % instruction-cache hit rate).  The round kind is warp-uniform.
GL_FN void permute(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], ROUND_ADD[i]);
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        if (full) {
#pragma unroll
            for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        } else {
            s[0] = sbox(s[0]);
        }
        mds_layer<false>(s, nullptr);
        if (r + 1 < 30) {
            const unsigned long long* nxt = &ROUND_ADD[(r + 1) * WIDTH];
            const bool next_full = (r + 1 < 4) || (r + 1 >= 26);
            if (next_full) {
#pragma unroll
                for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], nxt[i]);
            } else {
                s[0] = add_const(s[0], nxt[0]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::canon(s[i]);
}
#endif

}  // namespace poseidon
