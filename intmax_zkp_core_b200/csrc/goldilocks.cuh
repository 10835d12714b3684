// Goldilocks field arithmetic for sm_100a (p = 2^64 - 2^32 + 1).
//
// Replaces plonky2_field `GoldilocksField` (InternetMaximalism/plonky2 @ f99ed9c,
// field/src/goldilocks_field.rs; pinned by /root/reference/Cargo.toml:12) on the device.
// B200 has no 64-bit integer multiplier: every product is built from IMAD.WIDE.U32 (fma pipe) and the
// Solinas reduction 2^64 = 2^32 - 1, 2^96 = -1 from IADD3 carry chains (alu pipe).  Values travel as
// uint64_t; "canonical" means < p.  Functions say which operands must be canonical.
#pragma once
#include <cstdint>

// B200ZKP_HOST_EMU: tests/emu compiles the kernel bodies with g++ and steps "threads" in a loop to check
// index logic without a GPU.  It is a test build flavour only; the product library never defines it.
#ifdef B200ZKP_HOST_EMU
#define GL_FN static inline
#define GL_HD static inline
#define GL_MFN inline
#define GL_CONST_TABLE static const
#else
#define GL_FN __device__ __forceinline__
#define GL_HD __host__ __device__ __forceinline__      /* small pure helpers the host code shares (argument checks) */
#define GL_MFN __device__ __forceinline__
#define GL_CONST_TABLE static __device__ __constant__
#endif

namespace gl {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 P = 0xFFFFFFFF00000001ull;
static constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod p

// a >= p  <=>  hi == 0xFFFFFFFF and lo != 0, and then a - p = (0 : lo - 1)
GL_FN u64 canon(u64 a) {
    u32 lo = (u32)a, hi = (u32)(a >> 32);
    bool q = (hi == 0xFFFFFFFFu) & (lo != 0u);
    lo -= q ? 1u : 0u;
    hi = q ? 0u : hi;
    return ((u64)hi << 32) | lo;
}

#ifdef B200ZKP_HOST_EMU
// ---- plain C versions (host emulation build of the kernel bodies; same results as the PTX below)
// a, b canonical -> canonical
GL_FN u64 add(u64 a, u64 b) {
    u64 s = a + b;
    // a + b < 2p: wrapped past 2^64 (add EPS back == subtract p) or landed in [p, 2^64)
    return (s < a || s >= P) ? s - P : s;
}
// a, b canonical -> canonical
GL_FN u64 sub(u64 a, u64 b) {
    u64 d = a - b;
    return (a < b) ? d + P : d;
}
// (hi:lo) 128-bit -> u64 congruent mod p, NOT necessarily canonical.
GL_FN u64 reduce128(u64 lo, u64 hi) {
    u32 x2 = (u32)hi, x3 = (u32)(hi >> 32);
    u64 t0 = lo - x3;
    if (lo < (u64)x3) t0 -= EPS;          // borrow: -2^64 == -EPS
    u64 t1 = (u64)x2 * EPS;               // x2 * (2^32 - 1) < 2^64
    u64 r = t0 + t1;
    if (r < t1) r += EPS;                 // carry: +2^64 == +EPS (cannot carry twice)
    return r;
}
// a arbitrary u64, c canonical -> arbitrary u64 congruent to a + c
GL_FN u64 add_nc(u64 a, u64 c) {
    u64 s = a + c;
    return (s < a) ? s + EPS : s;         // single wrap: cannot wrap twice because c < p
}
// a arbitrary u64, c canonical -> arbitrary u64 congruent to a - c
GL_FN u64 sub_nc(u64 a, u64 c) {
    u64 d = a - c;
    return (a < c) ? d - EPS : d;         // a - c + 2^64 >= 2^64 - p + 1 = EPS: cannot borrow twice
}
// any u64 -> canonical (same as canon; the device version is a carry chain)
GL_FN u64 canon_cc(u64 a) { return a >= P ? a - P : a; }
#else
// ---- sm_100a versions: IADD3 carry chains written in PTX.  ptxas fuses `mul.wide.u32 + add.cc.u64` into
// one IMAD.WIDE.U32 with a carry-out predicate, and `subc 0,0` turns a carry/borrow into a 0 / 0xFFFFFFFF
// mask without compare+select.  A modular product is 14 SASS instructions (23 as compiled from C): 7 for the 128-bit
// product, 7 for the fold below.
// (hi:lo) 128-bit -> u64 congruent mod p, NOT necessarily canonical.
GL_FN u64 reduce128(u64 lo, u64 hi) {
    // r = lo - x3 + x2 * EPS lies in (-2^32, 2^65): one signed wrap count k = carry - borrow in {-1, 0, 1}, one fold of k * 2^64 = k * EPS
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, x2, x3, k, ks;\n\t"
        ".reg .u64 t, u;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "mov.b64 {x2, x3}, %2;\n\t"
        "sub.cc.u32 l0, l0, x3;\n\t"          // lo - x3      (2^96 == -1)
        "subc.cc.u32 l1, l1, 0;\n\t"
        "subc.u32 k, 0, 0;\n\t"               // k = -borrow
        "mov.b64 t, {l0, l1};\n\t"
        "mul.wide.u32 u, x2, 0xFFFFFFFF;\n\t" // x2 * EPS     (2^64 == EPS)
        "add.cc.u64 t, t, u;\n\t"
        "addc.u32 k, k, 0;\n\t"               // k += carry
        "shr.s32 ks, k, 31;\n\t"
        "mov.b64 {l0, l1}, t;\n\t"
        "sub.cc.u32 l0, l0, k;\n\t"           // + k * EPS == + k * 2^32 - k  (cannot wrap again)
        "subc.u32 l1, l1, ks;\n\t"
        "add.u32 l1, l1, k;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(lo), "l"(hi));
    return r;
}
// a arbitrary u64, c canonical -> arbitrary u64 congruent to a + c
GL_FN u64 add_nc(u64 a, u64 c) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, m;\n\t"
        ".reg .u64 t;\n\t"
        "add.cc.u64 t, %1, %2;\n\t"
        "addc.u32 m, 0, 0;\n\t"               // m = carry (0/1); NB: subc after add.cc has the wrong polarity
        "mov.b64 {l0, l1}, t;\n\t"
        "sub.cc.u32 l0, l0, m;\n\t"           // + m * EPS == + m * 2^32 - m (cannot carry twice)
        "subc.u32 l1, l1, 0;\n\t"
        "add.u32 l1, l1, m;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(a), "l"(c));
    return r;
}
// a arbitrary u64, c canonical -> arbitrary u64 congruent to a - c  (a - c + 2^64 >= EPS: one fold, cannot borrow twice).
// Same instructions as `sub`; the name states the weaker contract the lazy butterflies rely on.
GL_FN u64 sub_nc(u64 a, u64 c) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, m;\n\t"
        ".reg .u64 t;\n\t"
        "sub.cc.u64 t, %1, %2;\n\t"
        "subc.u32 m, 0, 0;\n\t"               // 0xFFFFFFFF on borrow: -2^64 == -EPS
        "mov.b64 {l0, l1}, t;\n\t"
        "sub.cc.u32 l0, l0, m;\n\t"
        "subc.u32 l1, l1, 0;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(a), "l"(c));
    return r;
}
// any u64 -> canonical: a + EPS carries out of 64 bits iff a >= p, and then the low 64 bits are a - p
// (4 SASS instructions: IADD3 + IADD3.X with a carry-out predicate + 2 SEL, against 6 for the compare form `canon`)
GL_FN u64 canon_cc(u64 a) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, s0, s1, m;\n\t"
        ".reg .pred q;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "add.cc.u32 s0, l0, 0xFFFFFFFF;\n\t"
        "addc.cc.u32 s1, l1, 0;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "setp.ne.u32 q, m, 0;\n\t"
        "selp.u32 l0, s0, l0, q;\n\t"
        "selp.u32 l1, s1, l1, q;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(a));
    return r;
}
// a canonical, b canonical -> canonical
GL_FN u64 sub(u64 a, u64 b) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 l0, l1, m;\n\t"
        ".reg .u64 t;\n\t"
        "sub.cc.u64 t, %1, %2;\n\t"
        "subc.u32 m, 0, 0;\n\t"               // 0xFFFFFFFF on borrow: add p == subtract EPS
        "mov.b64 {l0, l1}, t;\n\t"
        "sub.cc.u32 l0, l0, m;\n\t"
        "subc.u32 l1, l1, 0;\n\t"
        "mov.b64 %0, {l0, l1};\n\t"
        "}" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a, b canonical -> canonical:  a + b = a - (p - b)
GL_FN u64 add(u64 a, u64 b) { return sub(a, P - b); }
#endif
GL_FN u64 neg(u64 a) { return a ? P - a : 0; }

// a * b + c with 32-bit a, b and a 64-bit accumulator: exactly one IMAD.WIDE.U32 (fma pipe)
GL_FN u64 mad_wide(u32 a, u32 b, u64 c) {
#ifdef B200ZKP_HOST_EMU
    return (u64)a * b + c;
#else
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#endif
}

// any u64 operands -> congruent u64 (not necessarily canonical)
GL_FN u64 mul_nc(u64 a, u64 b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return reduce128((u64)p, (u64)(p >> 64));
}
// any u64 operands -> canonical
GL_FN u64 mul(u64 a, u64 b) { return canon(mul_nc(a, b)); }

#ifdef B200ZKP_HOST_EMU
GL_FN u32 brev32(u32 x) { u32 r = 0; for (int i = 0; i < 32; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
GL_FN u64 brev64(u64 x) { u64 r = 0; for (int i = 0; i < 64; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
GL_FN u64 ldg(const u64* p) { return *p; }
#else
GL_FN u32 brev32(u32 x) { return __brev(x); }
GL_FN u64 brev64(u64 x) { return __brevll(x); }
GL_FN u64 ldg(const u64* p) { return __ldg(p); }
#endif

}  // namespace gl
