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
GL_FN u64 combine3(u32 O0, u32 O1, u32 O2, u64 rc, u32 zero = 0u) {
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
        "}" : "=l"(r) : "r"(O0), "r"(O1), "r"(O2), "l"(rc), "r"(zero));
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
// which only the ALU pipe has.  kAluSp/kAluUv/kAluC/kAluNorm/kAluInj choose which groups of adds are pinned that way; the
// setting below balances the two pipes for the whole permutation (tools/sass_mix.py: 12.1 k slots on either pipe,
// down from 15.9 k on the multiplier pipe).
#ifdef B200ZKP_HOST_EMU
static const u32 OPAQUE_ZERO = 0;
#else
static __device__ __constant__ u32 OPAQUE_ZERO = 0;
#endif
// chosen on the GPU (profiles/r1c_poseidon_variants.log): pinned are the first butterfly level, the (1, 2, 1) product, the limb
// re-normalisation and the S-box injection; the second butterfly level and the two larger ring products stay with ptxas
static constexpr bool kAluSp = true, kAluUv = false, kAluC = true, kAluNorm = true, kAluInj = true;

// The state between two linear layers, in the split basis of Z[t] / (t^12 - 1) = (t^3 - 1)(t^3 + 1)(t^6 + 1): three limb
// planes of U[3], V[3], W[6] (signed 32-bit; tools/poseidon_crt_model.py bounds every intermediate by interval arithmetic).
struct SplitState {
    int U[3][3], V[3][3], W[3][6];
};

GL_FN u32 limb_of(u64 x, int L) {
    return (L == 0) ? ((u32)x & 0x3FFFFFu) : (L == 1) ? ((u32)(x >> 22) & 0x1FFFFFu) : (u32)(x >> 43);
}

// words -> split basis (the butterflies of the MDS layer); z8[L] = 8 * limb L of word 0 (the diag(8, 0, ..) term)
GL_FN void split_forward(const u64 (&s)[WIDTH], SplitState& c, int (&z8)[3], u32 Z) {
#pragma unroll
    for (int L = 0; L < 3; L++) {
        u32 l[WIDTH];
#pragma unroll
        for (int i = 0; i < WIDTH; i++) l[i] = limb_of(s[i], L);
        u32 sp[6];
#pragma unroll
        for (int i = 0; i < 6; i++) {
            sp[i] = l[i] + l[i + 6] + (kAluSp ? Z : 0u);
            c.W[L][i] = (int)(l[i] - l[i + 6] + (kAluSp ? Z : 0u));
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            c.U[L][i] = (int)(sp[i] + sp[i + 3] + (kAluUv ? Z : 0u));
            c.V[L][i] = (int)(sp[i] - sp[i + 3] + (kAluUv ? Z : 0u));
        }
        z8[L] = (int)(8u * l[0]);
    }
}

// three signed limbs of one word back to 22 / 21 / 21 bits (+ a small signed excess), same value mod p:
// carries upwards, then the part above 2^64 folds back through 2^64 = 2^32 - 1
GL_FN void normalise(int& a0, int& a1, int& a2, u32 Z) {
    const int zn = (int)(kAluNorm ? Z : 0u);
    const int c0 = a0 >> 22, n0 = a0 & 0x3FFFFF;
    const int t1 = a1 + c0 + zn;
    const int c1 = t1 >> 21, n1 = t1 & 0x1FFFFF;
    const int t2 = a2 + c1 + zn;
    const int top = t2 >> 21, n2 = t2 & 0x1FFFFF;
    a0 = n0 - top + zn;
    a1 = n1 + top * 1024;
    a2 = n2;
}

// the three ring products of one limb plane, unscaled:  Cq = U (*) (1, 2, 1),  D = V (*) (-1, -8, 2) mod t^3 + 1,
// B = W (*) Q mod t^6 + 1;  the layer is  (64 Cq, 4 D, 2 B)  in the split basis and  16 Cq +- D +- B  in words
GL_FN void ring_products(const int (&U)[3], const int (&V)[3], const int (&W)[6], int (&Cq)[3], int (&D)[3], int (&B)[6],
                         u32 Z) {
    const int T = U[0] + U[1] + U[2];
    Cq[0] = T + U[2] + (int)(kAluC ? Z : 0u);
    Cq[1] = T + U[0] + (int)(kAluC ? Z : 0u);
    Cq[2] = T + U[1] + (int)(kAluC ? Z : 0u);
    D[0] = 8 * V[2] - V[0] - 2 * V[1];
    D[1] = -8 * V[0] - V[1] - 2 * V[2];
    D[2] = 2 * V[0] - 8 * V[1] - V[2];
    constexpr int Qc[6] = {2, -4, 16, 1, -1, -1};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int acc = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) {
            const int q = (k >= i) ? Qc[k - i] : -Qc[k - i + 6];
            acc += W[i] * q;
        }
        B[k] = acc;
    }
}

// linear layer that stays in the split basis (the next round is a partial round), every word re-normalised
GL_FN void layer_stay(SplitState& c, const int (&z8)[3], u32 Z) {
#pragma unroll
    for (int L = 0; L < 3; L++) {
        int Cq[3], D[3], B[6];
        ring_products(c.U[L], c.V[L], c.W[L], Cq, D, B, Z);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            c.U[L][k] = 64 * Cq[k] + (k == 0 ? z8[L] : 0);
            c.V[L][k] = 4 * D[k] + (k == 0 ? z8[L] : 0);
        }
#pragma unroll
        for (int k = 0; k < 6; k++) c.W[L][k] = 2 * B[k] + (k == 0 ? z8[L] : 0);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        normalise(c.U[0][k], c.U[1][k], c.U[2][k], Z);
        normalise(c.V[0][k], c.V[1][k], c.V[2][k], Z);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) normalise(c.W[0][k], c.W[1][k], c.W[2][k], Z);
}

// linear layer back to words: out[k] = A[k] + B[k], out[k + 6] = A[k] - B[k] with A = 16 Cq +- D, plus the next round's
// constants `nxt` (they ride on the limb packing: 3 instructions per word instead of 6); `bias` (a multiple of
// the opaque zero, see above) lifts limbs that may be negative after partial rounds; the constants that follow absorb it
GL_FN void layer_leave(const SplitState& c, const int (&z8)[3], u64 (&s)[WIDTH], u32 bias, u32 Z, const unsigned long long* nxt) {
    u32 o[3][WIDTH];
#pragma unroll
    for (int L = 0; L < 3; L++) {
        int Cq[3], D[3], B[6], A[6];
        ring_products(c.U[L], c.V[L], c.W[L], Cq, D, B, Z);
#pragma unroll
        for (int k = 0; k < 3; k++) { A[k] = 16 * Cq[k] + D[k]; A[k + 3] = 16 * Cq[k] - D[k]; }
#pragma unroll
        for (int k = 0; k < 6; k++) {
            o[L][k] = (u32)(A[k] + B[k]) + bias;
            o[L][k + 6] = (u32)(A[k] - B[k]) + bias;
        }
        o[L][0] += (u32)z8[L];
    }
#pragma unroll
    for (int r = 0; r < WIDTH; r++) s[r] = combine3(o[0][r], o[1][r], o[2][r], (u64)nxt[r], Z);
}

// v / 4 mod p:  v = 4 q + r,  r / 4 = ((4 - r) << 62) - ((4 - r) << 30) + 1  (r = 0 gives p, which the carry fold removes)
GL_FN u64 div4(u64 v) {
    const u32 sh = (u32)v << 30;
    const u64 t = ((u64)(~sh) << 32) | (u64)(sh + 1u);
    return gl::add_nc(v >> 2, t);
}

// head of a partial round in the split basis: word 0 = (U0 + V0 + 2 W0) / 4 is packed, pushed through the S-box, and the
// difference is put back into the three components; cst = round constant - 2^21 (1 + 2^22 + 2^43) (SPLIT_ADD)
GL_FN void partial_head(SplitState& c, u64 cst, int (&z8)[3], u32 Z) {
    const int zi = (int)(kAluInj ? Z : 0u);
    u32 E[3];
#pragma unroll
    for (int L = 0; L < 3; L++) E[L] = (u32)(c.U[L][0] + c.V[L][0] + 2 * c.W[L][0] + (1 << 23));
    const u64 e = div4(combine3(E[0], E[1], E[2], 0ull, Z));
    constexpr int kDBias = 1 << 21;
    const u64 z = sbox(add_const(e, cst));
#pragma unroll
    for (int L = 0; L < 3; L++) {
        const int zl = (int)limb_of(z, L);
        const int d = zl - (int)limb_of(e, L) + kDBias;
        c.U[L][0] += d + zi;
        c.V[L][0] += d + zi;
        c.W[L][0] += d + zi;
        z8[L] = 8 * zl;
    }
}

// s + w * x  (all arbitrary u64) -> arbitrary u64
GL_FN u64 mul_add_nc(u64 w, u64 x, u64 s) {
    unsigned __int128 p = (unsigned __int128)w * x + s;   // <= (2^64-1)^2 + 2^64 - 1 < 2^128
    return gl::reduce128((u64)p, (u64)(p >> 64));
}

// In-place permutation; input words arbitrary u64, output words arbitrary u64 (congruent; `permute` canonicalises).
// One loop over all 30 rounds with a single copy of every block (S-box row, butterflies, ring products, packing): the
// whole permutation stays resident in the instruction cache (a two-loop form of 59 KB ran at a 67 % hit rate with "no
// instruction" as the top stall).  The round kind is warp-uniform.  Rounds 3..24 leave the state in the split basis
// (the next round is partial), every other round packs it back into words for the twelve S-boxes that follow.
GL_FN void permute_nc(u64 (&s)[WIDTH]) {
    using namespace poseidon_tables;
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = add_const(s[i], SPLIT_ADD[i]);
    SplitState c;
#pragma unroll
    for (int L = 0; L < 3; L++) {
#pragma unroll
        for (int k = 0; k < 3; k++) { c.U[L][k] = 0; c.V[L][k] = 0; }
#pragma unroll
        for (int k = 0; k < 6; k++) c.W[L][k] = 0;
    }
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const bool stay = (r >= 3) && (r < 25);
        const u32 Z = OPAQUE_ZERO;
        int z8[3];
        if (full) {
#pragma unroll
            for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
            split_forward(s, c, z8, Z);
            if (stay) {                     // r == 3: U holds sums of four limbs, too wide for the x256 of layer_stay
#pragma unroll
                for (int k = 0; k < 3; k++) normalise(c.U[0][k], c.U[1][k], c.U[2][k], Z);
            }
        } else {
            partial_head(c, SPLIT_ADD[r * WIDTH], z8, Z);
        }
        if (stay) {
            layer_stay(c, z8, Z);
        } else {
            // the constants of the next round ride on the limb packing (row 30 of SPLIT_ADD is all zero)
            layer_leave(c, z8, s, (full ? 0u : (1u << 30)) + Z, Z, &SPLIT_ADD[(r + 1) * WIDTH]);
        }
    }
}

// In-place permutation; input words arbitrary u64, output canonical.
GL_FN void permute(u64 (&s)[WIDTH]) {
    permute_nc(s);
#pragma unroll
    for (int i = 0; i < WIDTH; i++) s[i] = gl::canon(s[i]);
}

// permutation whose caller keeps only the digest words 0..3 (compressions, the last permutation of a sponge)
GL_FN void permute_digest(u64 (&s)[WIDTH]) { permute(s); }

}  // namespace poseidon
