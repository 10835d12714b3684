/* ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the product path
 * (intmax_zkp_core_b200/, include/, host/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use anything under oracle/.
 *
 * Goldilocks field, p = 2^64 - 2^32 + 1.  Restates plonky2_field `GoldilocksField`
 * (InternetMaximalism/plonky2 @ f99ed9c, field/src/goldilocks_field.rs — NOT present under
 * /root/reference, pinned by /root/reference/Cargo.toml:12 and Cargo.lock:333-335).
 * Unlike plonky2 the oracle keeps every value canonical (< p) at all times, for clarity.
 */
#ifndef ORACLE_GL_H
#define ORACLE_GL_H
#include <stdint.h>
#include <stddef.h>

typedef uint64_t gl_t;
typedef unsigned __int128 u128;
#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL            /* 2^64 mod p */
#define GL_GENERATOR 7ULL               /* MULTIPLICATIVE_GROUP_GENERATOR == coset_shift() */
#define GL_TWO_ADICITY 32
#define GL_POWER_OF_TWO_GENERATOR 1753635133440165772ULL /* 7^((p-1)/2^32), order 2^32 */

static inline gl_t gl_canon(gl_t a) { return a >= GL_P ? a - GL_P : a; }
static inline gl_t gl_add(gl_t a, gl_t b) { u128 s = (u128)a + b; return (gl_t)(s >= GL_P ? s - GL_P : s); }
static inline gl_t gl_sub(gl_t a, gl_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline gl_t gl_neg(gl_t a) { return a ? GL_P - a : 0; }
/* definitional product: 128-bit remainder.  Slow but beyond argument. */
static inline gl_t gl_mul_slow(gl_t a, gl_t b) { return (gl_t)(((u128)a * b) % GL_P); }
/* reduce128 as in plonky2 (2^64 = 2^32-1, 2^96 = -1), canonicalised at the end. */
static inline gl_t gl_reduce128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;                 /* borrow: -2^64 = -EPS */
    uint64_t t1 = hi_lo * GL_EPS;
    uint64_t r = t0 + t1;
    if (r < t1) r += GL_EPS;                      /* carry: +2^64 = +EPS */
    return gl_canon(r);
}
static inline gl_t gl_mul(gl_t a, gl_t b) { return gl_reduce128((u128)a * b); }
static inline gl_t gl_pow(gl_t b, uint64_t e) {
    gl_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, b); b = gl_mul(b, b); e >>= 1; }
    return r;
}
static inline gl_t gl_inv(gl_t a) { return gl_pow(a, GL_P - 2); }
/* primitive_root_of_unity(n_log) = g2^(2^(32-n_log)) */
static inline gl_t gl_root_of_unity(unsigned n_log) {
    gl_t w = GL_POWER_OF_TWO_GENERATOR;
    for (unsigned i = n_log; i < GL_TWO_ADICITY; i++) w = gl_mul(w, w);
    return w;
}
static inline uint64_t bitrev64(uint64_t x, unsigned bits) {
    uint64_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
#endif
