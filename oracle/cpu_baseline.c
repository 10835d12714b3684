/* ORACLE / CPU BASELINE — TEST AND BENCH INFRASTRUCTURE ONLY (see gl.h).
 *
 * Multithreaded CPU restatement of plonky2's commitment path with plonky2's own algorithm choices and
 * task partitioning (plonky2 @ f99ed9c; the Rust prover cannot be built here: no cargo/rustc, the crate
 * is an un-vendored git dependency of /root/reference/Cargo.toml:12).  It is what bench.py times as
 * `cpu_baseline` (kind "port") and under `--impl reference`; it is checked bit-for-bit against
 * oracle.c in tests/.
 *
 *   from_values:  one task per column: ifft = fft_classic (bit-reverse + radix-2 DIT over a precomputed
 *                 root table) + index reversal / n^-1 scaling            (field/src/fft.rs)
 *   lde_values:   one task per column: coeff_j * 7^j, zero-pad to N, fft_classic with
 *                 zero_factor = rate_bits (the first r layers only replicate)  (polynomial/mod.rs)
 *   transpose + reverse_index_bits_in_place: plonky2 does this serially (util/transpose); here it is
 *                 parallel over row blocks, i.e. the baseline is FASTER than plonky2 at this stage.
 *   MerkleTree::new: one task per cap subtree, recursive fork-join below (hash/merkle_tree.rs), Poseidon
 *                 with the fast partial rounds and u128 MDS accumulation (hash/poseidon.rs); plonky2 adds
 *                 hand-written AVX2/asm variants of the same arithmetic, this port relies on
 *                 -O3 -march=native.
 */
#include "gl.h"
#include "poseidon_fast_tables.h"
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#define WIDTH 12
static const uint64_t CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

/* non-canonical-tolerant helpers: values are arbitrary u64 congruent mod p, canonicalised on output.
 * Written the way plonky2's scalar x86-64 path is (branch-free reduce128, u128 dot products with a carry word,
 * MDS on 32-bit halves); plonky2 additionally ships hand-written asm / AVX2 variants of the same arithmetic. */
static inline uint64_t red128(u128 x) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    uint64_t hh = hi >> 32, hl = hi & GL_EPS;
    uint64_t t0 = lo - hh;
    t0 -= (uint64_t)(-(int64_t)(lo < hh)) & GL_EPS;       /* borrow: -2^64 = -EPS */
    uint64_t t1 = hl * GL_EPS;
    uint64_t r = t0 + t1;
    r += (uint64_t)(-(int64_t)(r < t1)) & GL_EPS;         /* carry: +2^64 = +EPS */
    return r;
}
static inline uint64_t mulnc(uint64_t a, uint64_t b) { return red128((u128)a * b); }
static inline uint64_t sbox(uint64_t x) {
    uint64_t x2 = mulnc(x, x), x4 = mulnc(x2, x2), x3 = mulnc(x2, x);
    return mulnc(x3, x4);
}
/* sum of up to 16 u64*u64 products: 128-bit accumulator + carry word; 2^128 = -2^32 (mod p) */
typedef struct { u128 acc; uint64_t over; } acc160_t;
static inline void acc_mac(acc160_t* a, uint64_t x, uint64_t y) {
    u128 p = (u128)x * y;
    a->acc += p;
    a->over += a->acc < p;
}
static inline uint64_t acc_reduce(const acc160_t* a) {
    uint64_t r = gl_canon(red128(a->acc));
    return gl_sub(r, a->over << 32);
}
static inline void mds(uint64_t s[WIDTH], const uint64_t* addc) {
    /* split into 32-bit halves so the 12x12 small-constant products accumulate in plain u64 (plonky2's
     * mds_row_shf works on the same idea); doubled array avoids the index wrap */
    uint64_t lo[2 * WIDTH], hi[2 * WIDTH];
    for (int i = 0; i < WIDTH; i++) { lo[i] = lo[i + WIDTH] = (uint32_t)s[i]; hi[i] = hi[i + WIDTH] = s[i] >> 32; }
#pragma GCC unroll 12
    for (int r = 0; r < WIDTH; r++) {
        uint64_t L = 0, H = 0;
#pragma GCC unroll 12
        for (int i = 0; i < WIDTH; i++) { L += lo[i + r] * CIRC[i]; H += hi[i + r] * CIRC[i]; }
        if (r == 0) { L += lo[0] * 8; H += hi[0] * 8; }
        u128 acc = (u128)L + ((u128)H << 32);
        if (addc) acc += addc[r];
        s[r] = red128(acc);
    }
}
void cpub_permute(uint64_t s[WIDTH]) {
    for (int i = 0; i < WIDTH; i++) s[i] = red128((u128)s[i] + RC_FULL[i]);
    for (int r = 0; r < 4; r++) {
#pragma GCC unroll 12
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        mds(s, r < 3 ? &RC_FULL[(r + 1) * WIDTH] : FAST_FIRST);
    }
    {
        uint64_t t[WIDTH - 1];
        for (int r = 0; r < WIDTH - 1; r++) {
            acc160_t a = {0, 0};
            for (int c = 0; c < WIDTH - 1; c++) acc_mac(&a, FAST_INIT[r * (WIDTH - 1) + c], s[1 + c]);
            t[r] = acc_reduce(&a);
        }
        memcpy(s + 1, t, sizeof t);
    }
    for (int i = 0; i < 22; i++) {
        uint64_t s0 = red128((u128)sbox(s[0]) + FAST_POST[i]);
        acc160_t a = {(u128)s0 * 25, 0};
#pragma GCC unroll 11
        for (int j = 0; j < WIDTH - 1; j++) acc_mac(&a, FAST_VHAT[i * (WIDTH - 1) + j], s[1 + j]);
#pragma GCC unroll 11
        for (int j = 0; j < WIDTH - 1; j++) s[1 + j] = red128((u128)FAST_WHAT[i * (WIDTH - 1) + j] * s0 + s[1 + j]);
        s[0] = acc_reduce(&a);
    }
    for (int i = 0; i < WIDTH; i++) s[i] = red128((u128)s[i] + RC_FULL[4 * WIDTH + i]);
    for (int r = 4; r < 8; r++) {
#pragma GCC unroll 12
        for (int i = 0; i < WIDTH; i++) s[i] = sbox(s[i]);
        mds(s, r < 7 ? &RC_FULL[(r + 1) * WIDTH] : NULL);
    }
    for (int i = 0; i < WIDTH; i++) s[i] = gl_canon(s[i]);
}
static void hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]) {
    if (len <= 4) { for (size_t i = 0; i < 4; i++) out[i] = i < len ? gl_canon(in[i]) : 0; return; }
    uint64_t st[WIDTH] = {0};
    for (size_t off = 0; off < len; off += 8) {
        size_t c = len - off < 8 ? len - off : 8;
        memcpy(st, in + off, c * sizeof(uint64_t));
        cpub_permute(st);
    }
    memcpy(out, st, 32);
}
static void two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    uint64_t st[WIDTH] = {0};
    memcpy(st, l, 32); memcpy(st + 4, r, 32);
    cpub_permute(st);
    memcpy(out, st, 32);
}
/* hashes `rows` row-major leaves of `len` elements with every thread; returns seconds (bench sampling) */
double cpub_leaf_hash_rows(const uint64_t* leaves, uint64_t rows, uint32_t len, uint64_t* digests) {
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)rows; i++) hash_or_noop(leaves + (size_t)i * len, len, digests + 4 * (size_t)i);
    return omp_get_wtime() - t0;
}

/* ---- fft_classic over a precomputed root table (fft_root_table) ---- */
typedef struct { unsigned lg; gl_t** layer; } root_table_t;
static root_table_t make_root_table(unsigned lg) {
    root_table_t t; t.lg = lg; t.layer = (gl_t**)calloc(lg + 1, sizeof(gl_t*));
    for (unsigned lg_m = 1; lg_m <= lg; lg_m++) {
        size_t half = (size_t)1 << (lg_m - 1);
        t.layer[lg_m] = (gl_t*)malloc(half * sizeof(gl_t));
        gl_t w = gl_root_of_unity(lg_m), x = 1;
        for (size_t j = 0; j < half; j++) { t.layer[lg_m][j] = x; x = gl_mul(x, w); }
    }
    return t;
}
static void free_root_table(root_table_t* t) {
    for (unsigned i = 1; i <= t->lg; i++) free(t->layer[i]);
    free(t->layer);
}
static void bitrev_in_place(gl_t* v, unsigned lg) {
    size_t n = (size_t)1 << lg;
    for (size_t i = 0; i < n; i++) { size_t j = bitrev64(i, lg); if (i < j) { gl_t t = v[i]; v[i] = v[j]; v[j] = t; } }
}
/* r = zero_factor: the input had only its first n >> r entries non-zero before bit reversal */
static void fft_classic(gl_t* v, unsigned lg, unsigned r, const root_table_t* rt) {
    size_t n = (size_t)1 << lg;
    bitrev_in_place(v, lg);
    if (r > 0) { /* the first r layers of butterflies (a, 0) -> (a, a): replicate */
        size_t m = (size_t)1 << r;
        for (size_t k = 0; k < n; k += m) for (size_t j = 1; j < m; j++) v[k + j] = v[k];
    }
    for (unsigned lg_m = r + 1; lg_m <= lg; lg_m++) {
        size_t m = (size_t)1 << lg_m, half = m >> 1;
        const gl_t* w = rt->layer[lg_m];
        for (size_t k = 0; k < n; k += m)
            for (size_t j = 0; j < half; j++) {
                gl_t t = gl_mul(w[j], v[k + j + half]), u = v[k + j];
                v[k + j] = gl_add(u, t);
                v[k + j + half] = gl_sub(u, t);
            }
    }
}

static void fill_subtree(gl_t* digests, size_t digests_len, const uint64_t* leaves, size_t n_leaves,
                         size_t leaf_len, gl_t out[4]) {
    if (digests_len == 0) { hash_or_noop(leaves, leaf_len, out); return; }
    size_t half = digests_len / 2;
    gl_t l[4], r[4];
    if (n_leaves >= 1024) {
#pragma omp task shared(l)
        fill_subtree(digests, half - 1, leaves, n_leaves / 2, leaf_len, l);
#pragma omp task shared(r)
        fill_subtree(digests + 4 * (half + 1), half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
#pragma omp taskwait
    } else {
        fill_subtree(digests, half - 1, leaves, n_leaves / 2, leaf_len, l);
        fill_subtree(digests + 4 * (half + 1), half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
    }
    memcpy(digests + 4 * (half - 1), l, 32);
    memcpy(digests + 4 * half, r, 32);
    two_to_one(l, r, out);
}

int cpub_threads(void) { return omp_get_max_threads(); }
/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the bench sets the team size explicitly */
void cpub_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* times[5] = { ifft, lde, transpose+bitrev, merkle, total } seconds */
int cpub_commit(const uint64_t* in, int is_coeffs, uint32_t n_log, uint32_t k, uint32_t rate_bits,
                uint32_t cap_height, uint64_t* coeffs, uint64_t* leaves, uint64_t* digests, uint64_t* cap,
                double times[5]) {
    size_t n = (size_t)1 << n_log, N = n << rate_bits;
    unsigned N_log = n_log + rate_bits;
    if (cap_height > N_log) return -1;
    root_table_t rt = make_root_table(N_log); /* built once in CircuitBuilder::build, not part of a commit */
    uint64_t* lde = (uint64_t*)malloc((size_t)k * N * sizeof(uint64_t));
    if (!lde) { free_root_table(&rt); return -2; }
    gl_t n_inv = gl_inv((gl_t)(n % GL_P));
    double t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic)
    for (int64_t c = 0; c < (int64_t)k; c++) {
        gl_t* cf = coeffs + (size_t)c * n;
        for (size_t i = 0; i < n; i++) cf[i] = gl_canon(in[(size_t)c * n + i]);
        if (!is_coeffs && n > 1) {
            fft_classic(cf, n_log, 0, &rt);
            cf[0] = gl_mul(cf[0], n_inv);
            cf[n / 2] = gl_mul(cf[n / 2], n_inv);
            for (size_t i = 1; i < n / 2; i++) {
                size_t j = n - i;
                gl_t ci = gl_mul(cf[j], n_inv), cj = gl_mul(cf[i], n_inv);
                cf[i] = ci; cf[j] = cj;
            }
        }
    }
    double t1 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic)
    for (int64_t c = 0; c < (int64_t)k; c++) {
        gl_t* v = lde + (size_t)c * N;
        const gl_t* cf = coeffs + (size_t)c * n;
        gl_t s = 1;
        for (size_t i = 0; i < n; i++) { v[i] = gl_mul(cf[i], s); s = gl_mul(s, GL_GENERATOR); }
        memset(v + n, 0, (N - n) * sizeof(gl_t));
        fft_classic(v, N_log, rate_bits, &rt);
    }
    double t2 = omp_get_wtime();
    /* transpose + reverse_index_bits_in_place (parallel here, serial in plonky2) */
#pragma omp parallel for schedule(static)
    for (int64_t jb = 0; jb < (int64_t)N; jb += 64) {
        size_t jend = (size_t)jb + 64 < N ? (size_t)jb + 64 : N;
        for (size_t j = (size_t)jb; j < jend; j++) {
            size_t src = bitrev64(j, N_log);
            uint64_t* row = leaves + j * k;
            for (uint32_t c = 0; c < k; c++) row[c] = lde[(size_t)c * N + src];
        }
    }
    double t3 = omp_get_wtime();
    free(lde);
    size_t n_cap = (size_t)1 << cap_height;
    size_t sub_leaves = N >> cap_height;
    size_t sub_digests = 2 * (sub_leaves - 1);
#pragma omp parallel
#pragma omp single
    for (size_t s = 0; s < n_cap; s++) {
#pragma omp task firstprivate(s)
        fill_subtree(digests + 4 * s * sub_digests, sub_digests, leaves + s * sub_leaves * k, sub_leaves, k, cap + 4 * s);
    }
    double t4 = omp_get_wtime();
    free_root_table(&rt);
    if (times) { times[0] = t1 - t0; times[1] = t2 - t1; times[2] = t3 - t2; times[3] = t4 - t3; times[4] = t4 - t0; }
    return 0;
}
