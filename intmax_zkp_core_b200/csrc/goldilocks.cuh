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
#define GL_MFN inline
#define GL_CONST_TABLE static const
#else
#define GL_FN __device__ __forceinline__
#define GL_MFN __device__ __forceinline__
#define GL_CONST_TABLE static __device__ __constant__
#endif

namespace gl {

typedef unsigned long long u64;
typedef unsigned int u32;

static constexpr u64 P = 0xFFFFFFFF00000001ull;
static constexpr u64 EPS = 0xFFFFFFFFull;  // 2^64 mod p

GL_FN u64 canon(u64 a) { return a >= P ? a - P : a; }

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
GL_FN u64 neg(u64 a) { return a ? P - a : 0; }

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
// any u64 operands -> congruent u64 (not necessarily canonical)
GL_FN u64 mul_nc(u64 a, u64 b) {
    unsigned __int128 p = (unsigned __int128)a * b;
    return reduce128((u64)p, (u64)(p >> 64));
}
// any u64 operands -> canonical
GL_FN u64 mul(u64 a, u64 b) { return canon(mul_nc(a, b)); }

GL_FN u64 sqr_nc(u64 a) { return mul_nc(a, a); }

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
