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
//   * S-box: 4 modular products of 16 SASS instructions each (4 IMAD.WIDE.U32 + carry chain + Solinas fold).
//   * MDS: every word is cut into 22 + 21 + 21-bit limbs so that all 12-term sums fit 32 bits; the 12x12 circulant is
//     split by x^12 - 1 = (x^3 - 1)(x^3 + 1)(x^6 + 1) into pieces whose constants are all +-2^j, so the layer is adds and
//     shift-adds only, spread evenly over the ALU and the multiplier pipe; limbs are recombined with one fold.
//   * partial rounds use the "pushed constant" form: the 11 idle words of every partial-round constant are pushed
//     forward through the MDS, so a partial round is  s0 = (s0 + c)^7 ; s = MDS s  with the same MDS body.
//   * one loop over all 30 rounds (single copy of the S-box row and of the MDS): ~25 KB of SASS, instruction-cache
//     resident.  tools/poseidon_derive.py derives and checks the constant schedule (ROUND_ADD).
// Words between rounds are arbitrary u64 (not canonical); `permute` canonicalises its output.
#pragma once
#include "goldilocks.cuh"
#include "poseidon_tables.cuh"

namespace poseidon {

using gl::u32;
using gl::u64;

static constexpr int WIDTH = 12;
static constexpr int RATE = 8;

GL_FN u64 sbox(u64 x) {
    u64 x2 = gl::mul_nc(x, x);
    u64 x4 = gl::mul_nc(x2, x2);
    u64 x3 = gl::mul_nc(x2, x);
    return gl::mul_nc(x3, x4);
}

// a arbitrary u64, c canonical constant -> arbitrary u64 congruent to a + c
GL_FN u64 add_const(u64 a, u64 c) { return gl::add_nc(a, c); }

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

// MDS on three limbs (22 + 21 + 21 bits) so that every term and every 12-term sum fits 32 bits, and without a single
// multiplication.  The 12x12 circulant is a product in Z[x] / (x^12 - 1); x^12 - 1 = (x^6 - 1)(x^6 + 1) splits it into a
// cyclic and a negacyclic 6x6 product with halved constants
//   P = (15, 24, 18, 17, 40, 14)   and   Q = (2, -4, 16, 1, -1, -1),
// and x^6 - 1 = (x^3 - 1)(x^3 + 1) splits the cyclic one again into 3x3 products with
//   P+ = (16, 32, 16) = 16 * (1, 2, 1)   and   P- = (-1, -8, 2).
// Every constant is +-2^j (this is how the matrix was chosen: plonky2's MDS_FREQ_BLOCK_* hold the same numbers in the
// frequency domain), so a layer is 71 adds / shift-adds per limb instead of 72 multiply-adds + 36 adds.
// out[k] = A[k] + B[k], out[k+6] = A[k] - B[k]; arithmetic is mod 2^32 (exact: true sums < 2^31).
//
// Pipe steering.  B200 issues IMAD* on the multiplier pipe and IADD3 / LOP3 / SHF / LEA on the ALU pipe, one warp
// instruction per two cycles each, and an IMAD.WIDE holds the multiplier pipe twice as long.  ptxas turns every two-input
// add into IMAD.IADD, which left the multiplier pipe 92 % busy with the ALU half idle (profiles/README.md).  A third
// operand that ptxas cannot prove to be zero (OPAQUE_ZERO, a __constant__ word) makes the add a three-input IADD3,
// which only the ALU pipe has.  kAluSp/kAluUv/kAluC/kAluOut choose which groups of adds are pinned that way; the
// setting below balances the two pipes for the whole permutation (tools/sass_mix.py: 12.1 k slots on either pipe,
// down from 15.9 k on the multiplier pipe).
#ifdef B200ZKP_HOST_EMU
static const u32 OPAQUE_ZERO = 0;
#else
static __device__ __constant__ u32 OPAQUE_ZERO = 0;
#endif
#ifndef B200ZKP_MDS_ALU_MASK
#define B200ZKP_MDS_ALU_MASK 13   // tuning builds only (tools/bench_variants.sh)
#endif
static constexpr bool kAluSp = B200ZKP_MDS_ALU_MASK & 1, kAluUv = B200ZKP_MDS_ALU_MASK & 2, kAluC = B200ZKP_MDS_ALU_MASK & 4,
                      kAluOut = B200ZKP_MDS_ALU_MASK & 8;

template <bool kAddConst>
GL_FN void mds_layer(u64 (&s)[WIDTH], const unsigned long long* addc) {
    u32 o[3][WIDTH];
    const u32 Z = OPAQUE_ZERO;
#pragma unroll
    for (int L = 0; L < 3; L++) {
        u32 l[WIDTH];
#pragma unroll
        for (int i = 0; i < WIDTH; i++)
            l[i] = (L == 0) ? ((u32)s[i] & 0x3FFFFFu) : (L == 1) ? ((u32)(s[i] >> 22) & 0x1FFFFFu) : (u32)(s[i] >> 43);
        u32 sp[6], sm[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            sp[i] = l[i] + l[i + 6] + (kAluSp ? Z : 0u);
            sm[i] = l[i] - l[i + 6] + (kAluSp ? Z : 0u);
        }
        // cyclic half: A = sp (*) P mod (x^6 - 1), through (x^3 - 1)(x^3 + 1)
        u32 u[3], v[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            u[i] = sp[i] + sp[i + 3] + (kAluUv ? Z : 0u);
            v[i] = sp[i] - sp[i + 3] + (kAluUv ? Z : 0u);
        }
        const u32 T = u[0] + u[1] + u[2];
        u32 Cq[3], D[3], A[6];                    // C = 16 * Cq = u (*) (16, 32, 16) mod (x^3 - 1)
        Cq[0] = T + u[2] + (kAluC ? Z : 0u);
        Cq[1] = T + u[0] + (kAluC ? Z : 0u);
        Cq[2] = T + u[1] + (kAluC ? Z : 0u);
        D[0] = 8u * v[2] - v[0] - 2u * v[1];      // D = v (*) (-1, -8, 2) mod (x^3 + 1)
        D[1] = 0u - 8u * v[0] - v[1] - 2u * v[2];
        D[2] = 2u * v[0] - 8u * v[1] - v[2];
#pragma unroll
        for (int k = 0; k < 3; k++) { A[k] = 16u * Cq[k] + D[k]; A[k + 3] = 16u * Cq[k] - D[k]; }
        // negacyclic half: B = sm (*) Q mod (x^6 + 1)
        constexpr int Qc[6] = {2, -4, 16, 1, -1, -1};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            u32 B = 0;
#pragma unroll
            for (int i = 0; i < 6; i++) {
                const int q = (k >= i) ? Qc[k - i] : -Qc[k - i + 6];
                B += sm[i] * (u32)q;
            }
            o[L][k] = A[k] + B + (kAluOut ? Z : 0u);
            o[L][k + 6] = A[k] - B + (kAluOut ? Z : 0u);
        }
        o[L][0] += 8u * l[0];     // diag(8, 0, ..., 0)
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) s[r] = combine3(o[0][r], o[1][r], o[2][r], kAddConst ? (u64)addc[r] : 0ull);
}

// s + w * x  (all arbitrary u64) -> arbitrary u64
GL_FN u64 mul_add_nc(u64 w, u64 x, u64 s) {
    unsigned __int128 p = (unsigned __int128)w * x + s;   // <= (2^64-1)^2 + 2^64 - 1 < 2^128
    return gl::reduce128((u64)p, (u64)(p >> 64));
}

// In-place permutation; input words arbitrary u64, output canonical.
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

}  // namespace poseidon
