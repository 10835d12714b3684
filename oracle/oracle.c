/* ORACLE — TEST INFRASTRUCTURE ONLY (see gl.h).  Single-threaded, clarity over speed.
 *
 * CPU restatement of the polynomial-commitment hot path of plonky2 @ f99ed9c
 * (InternetMaximalism/plonky2, pinned by /root/reference/Cargo.toml:12, Cargo.lock:333-335).
 * The plonky2 source is NOT under /root/reference and no Rust toolchain exists here, so every
 * function below restates the published algorithm and is
 * anchored on the reference's own call sites and fixtures:
 *
 *   Poseidon permutation / two_to_one / hash_no_pad / hash_pad  — PINNED by reference fixtures:
 *     /root/reference/src/transaction/circuits/mod.rs:211-218 (two_to_one(0,0)),
 *     /root/reference/src/rollup/circuits/mod.rs:104 (32-level zero-hash chain + three roots),
 *     /root/reference/src/bin/block_circuit.rs:81-88,157-164 + test_cases/block1_info.json
 *       (H(sk,sk) addresses, SMT transaction hashes),
 *     /root/reference/src/sparse_merkle_tree/goldilocks_poseidon/mod.rs:161-183 (hash_pad use),
 *     /root/reference/src/sparse_merkle_tree/gadgets/common.rs:87-101 (padding rule).
 *   NTT / coset LDE / leaf order / Merkle cap / digest layout — PARITY UNPINNED: no test or fixture
 *     in the reference holds an NTT output, an LDE row, a MerkleCap or proof bytes.
 *     Conventions restated from plonky2 (field/src/fft.rs, field/src/polynomial/mod.rs,
 *     plonky2/src/fri/oracle.rs, plonky2/src/hash/merkle_tree.rs) and cross-checked by an O(n^2)
 *     definitional DFT, an independent pure-Python twin (oracle/pyref.py) and SURVEY.md App. C.
 */
#include "gl.h"
#include <stdlib.h>
#include <string.h>

#define WIDTH 12
#define RATE 8
#define N_FULL_HALF 4
#define N_PARTIAL 22
#define N_ROUNDS 30

/* ------------------------------------------------------------------ Poseidon round constants
 * plonky2 poseidon_goldilocks.rs ALL_ROUND_CONSTANTS were produced by
 *   ChaCha8Rng::seed_from_u64(0); 360 x rng.gen_range(0..p)          
 * regenerated here at run time (rand_core seed_from_u64 = PCG32 stream; rand 0.8 single-sample
 * uniform u64).  tests/ compare against SURVEY.md App. A rows and tools/poseidon_derive.py. */
static uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static void chacha8_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) st[4 + i] = key[i];
    st[12] = (uint32_t)counter; st[13] = (uint32_t)(counter >> 32); st[14] = 0; st[15] = 0;
    uint32_t w[16];
    memcpy(w, st, sizeof w);
#define QR(a, b, c, d)                                                                      \
    w[a] += w[b]; w[d] = rotl32(w[d] ^ w[a], 16); w[c] += w[d]; w[b] = rotl32(w[b] ^ w[c], 12); \
    w[a] += w[b]; w[d] = rotl32(w[d] ^ w[a], 8);  w[c] += w[d]; w[b] = rotl32(w[b] ^ w[c], 7);
    for (int r = 0; r < 4; r++) {
        QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15)
        QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) out[i] = w[i] + st[i];
}

static gl_t g_rc[N_ROUNDS * WIDTH];
static int g_rc_ready = 0;
static void init_round_constants(void) {
    if (g_rc_ready) return;
    uint64_t s = 0;
    uint32_t key[8];
    for (int i = 0; i < 8; i++) {
        s = s * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27);
        uint32_t rot = (uint32_t)(s >> 59);
        key[i] = (x >> rot) | (x << ((32 - rot) & 31));
    }
    uint32_t buf[16];
    int pos = 16;
    uint64_t ctr = 0;
    int n = 0;
    while (n < N_ROUNDS * WIDTH) {
        uint32_t w2[2];
        for (int j = 0; j < 2; j++) {
            if (pos == 16) { chacha8_block(key, ctr++, buf); pos = 0; }
            w2[j] = buf[pos++];
        }
        uint64_t v = (uint64_t)w2[0] | ((uint64_t)w2[1] << 32);
        u128 prod = (u128)v * GL_P;
        if ((uint64_t)prod <= GL_P - 1) g_rc[n++] = (uint64_t)(prod >> 64);
    }
    g_rc_ready = 1;
}
void orc_round_constants(uint64_t out[N_ROUNDS * WIDTH]) {
    init_round_constants();
    memcpy(out, g_rc, sizeof g_rc);
}

/* ------------------------------------------------------------------ Poseidon permutation (A9)
 * plonky2/src/hash/poseidon.rs `Poseidon::poseidon`: 4 full + 22 partial + 4 full rounds, x^7,
 * MDS out[r] = sum_i in[(i+r)%12]*CIRC[i] + in[r]*DIAG[r].  Naive form (plonky2's fast partial
 * rounds are algebraically identical). */
static const uint64_t MDS_CIRC[WIDTH] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
static const uint64_t MDS_DIAG[WIDTH] = {8, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

static gl_t sbox7(gl_t x) {
    gl_t x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x2, x);
    return gl_mul(x3, x4);
}
static void mds_layer(gl_t s[WIDTH]) {
    gl_t o[WIDTH];
    for (int r = 0; r < WIDTH; r++) {
        u128 acc = 0; /* 12 * 41 * 2^64 < 2^74 */
        for (int i = 0; i < WIDTH; i++) acc += (u128)s[(i + r) % WIDTH] * MDS_CIRC[i];
        acc += (u128)s[r] * MDS_DIAG[r];
        o[r] = (gl_t)(acc % GL_P);
    }
    memcpy(s, o, sizeof o);
}
void orc_permute(uint64_t s[WIDTH]) {
    init_round_constants();
    for (int i = 0; i < WIDTH; i++) s[i] = gl_canon(s[i]);
    for (int r = 0; r < N_ROUNDS; r++) {
        for (int i = 0; i < WIDTH; i++) s[i] = gl_add(s[i], g_rc[r * WIDTH + i]);
        if (r < N_FULL_HALF || r >= N_FULL_HALF + N_PARTIAL)
            for (int i = 0; i < WIDTH; i++) s[i] = sbox7(s[i]);
        else
            s[0] = sbox7(s[0]);
        mds_layer(s);
    }
}

/* plonky2/src/hash/hashing.rs `hash_n_to_m_no_pad` with 4 outputs (A6): overwrite-mode sponge,
 * rate 8; a short last chunk overwrites only its own length. */
void orc_hash_no_pad(const uint64_t* in, size_t len, uint64_t out[4]) {
    gl_t st[WIDTH] = {0};
    for (size_t off = 0; off < len; off += RATE) {
        size_t c = len - off < RATE ? len - off : RATE;
        for (size_t i = 0; i < c; i++) st[i] = gl_canon(in[off + i]);
        orc_permute(st);
    }
    memcpy(out, st, 4 * sizeof(gl_t));
}
/* plonky2 `hash_pad` (hashing.rs `hash_n_to_m_with_pad`): append 1, zeros to a multiple of the
 * WIDTH (12, not the rate), last element 1.  Pinned by
 * /root/reference/src/sparse_merkle_tree/gadgets/common.rs:87-101 ([k,v,1,1,0,1]). */
void orc_hash_pad(const uint64_t* in, size_t len, uint64_t out[4]) {
    size_t padded = len + 2;
    while (padded % WIDTH) padded++;
    uint64_t* buf = (uint64_t*)calloc(padded, sizeof(uint64_t));
    memcpy(buf, in, len * sizeof(uint64_t));
    buf[len] = 1;
    buf[padded - 1] = 1;
    orc_hash_no_pad(buf, padded, out);
    free(buf);
}
/* `Hasher::hash_or_noop`: inputs of <= 4 elements are the digest, zero padded (A6). */
void orc_hash_or_noop(const uint64_t* in, size_t len, uint64_t out[4]) {
    if (len <= 4) {
        for (size_t i = 0; i < 4; i++) out[i] = i < len ? gl_canon(in[i]) : 0;
    } else {
        orc_hash_no_pad(in, len, out);
    }
}
/* `PoseidonHash::two_to_one` = hashing.rs `compress` (A8). */
void orc_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]) {
    gl_t st[WIDTH] = {0};
    for (int i = 0; i < 4; i++) { st[i] = gl_canon(l[i]); st[4 + i] = gl_canon(r[i]); }
    orc_permute(st);
    memcpy(out, st, 4 * sizeof(gl_t));
}

/* ------------------------------------------------------------------ NTT (A2, A4)
 * Definitional transform: out[i] = sum_j in[j] * w_n^(i*j), w_n = g2^(2^(32-n_log)). */
void orc_dft(const uint64_t* in, uint64_t* out, unsigned n_log) {
    size_t n = (size_t)1 << n_log;
    gl_t w = gl_root_of_unity(n_log);
    gl_t wi = 1;
    for (size_t i = 0; i < n; i++) {
        gl_t acc = 0, x = 1;
        for (size_t j = 0; j < n; j++) { acc = gl_add(acc, gl_mul(gl_canon(in[j]), x)); x = gl_mul(x, wi); }
        out[i] = acc;
        wi = gl_mul(wi, w);
    }
}
/* field/src/fft.rs `fft_classic`: reverse_index_bits_in_place then radix-2 DIT layers with
 * root_table[lg_m-1][j] = w_{2^lg_m}^j.  Natural order in and out. */
void orc_fft(uint64_t* v, unsigned n_log) {
    size_t n = (size_t)1 << n_log;
    for (size_t i = 0; i < n; i++) {
        v[i] = gl_canon(v[i]);
    }
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev64(i, n_log);
        if (i < j) { gl_t t = v[i]; v[i] = v[j]; v[j] = t; }
    }
    for (unsigned lg_m = 1; lg_m <= n_log; lg_m++) {
        size_t m = (size_t)1 << lg_m, half = m >> 1;
        gl_t wm = gl_root_of_unity(lg_m);
        for (size_t k = 0; k < n; k += m) {
            gl_t w = 1;
            for (size_t j = 0; j < half; j++) {
                gl_t t = gl_mul(w, v[k + j + half]), u = v[k + j];
                v[k + j] = gl_add(u, t);
                v[k + j + half] = gl_sub(u, t);
                w = gl_mul(w, wm);
            }
        }
    }
}
/* field/src/fft.rs `ifft_with_options`: forward transform, then scale by n^-1 and reverse all
 * entries but the first. */
void orc_ifft(uint64_t* v, unsigned n_log) {
    size_t n = (size_t)1 << n_log;
    orc_fft(v, n_log);
    gl_t n_inv = gl_inv((gl_t)(n % GL_P));
    if (n == 1) return;
    v[0] = gl_mul(v[0], n_inv);
    v[n / 2] = gl_mul(v[n / 2], n_inv);
    for (size_t i = 1; i < n / 2; i++) {
        size_t j = n - i;
        gl_t ci = gl_mul(v[j], n_inv), cj = gl_mul(v[i], n_inv);
        v[i] = ci; v[j] = cj;
    }
}
/* polynomial/mod.rs `lde(rate_bits)` + `coset_fft_with_options(shift = 7, ..)`:
 * out[i] = p(7 * w_N^i), i natural, N = n << rate_bits (A4). */
void orc_coset_lde(const uint64_t* coeffs, unsigned n_log, unsigned rate_bits, uint64_t* out) {
    size_t n = (size_t)1 << n_log, N = n << rate_bits;
    gl_t s = 1;
    for (size_t i = 0; i < N; i++) {
        if (i < n) { out[i] = gl_mul(gl_canon(coeffs[i]), s); s = gl_mul(s, GL_GENERATOR); }
        else out[i] = 0;
    }
    orc_fft(out, n_log + rate_bits);
}
/* slow second opinion: Horner evaluation of p at 7*w_N^i for one i */
uint64_t orc_eval_at_lde_point(const uint64_t* coeffs, unsigned n_log, unsigned rate_bits, uint64_t i) {
    size_t n = (size_t)1 << n_log;
    gl_t x = gl_mul(GL_GENERATOR, gl_pow(gl_root_of_unity(n_log + rate_bits), i));
    gl_t acc = 0;
    for (size_t j = n; j-- > 0;) acc = gl_add(gl_mul(acc, x), gl_canon(coeffs[j]));
    return acc;
}

/* ------------------------------------------------------------------ MerkleTree (A7)
 * plonky2/src/hash/merkle_tree.rs `fill_subtree` / `fill_digests_buf`: layout
 * left_subtree || left_child || right_child || right_subtree; roots only in the cap. */
static void fill_subtree(gl_t* digests /*4*len*/, size_t digests_len, const uint64_t* leaves,
                         size_t n_leaves, size_t leaf_len, gl_t out[4]) {
    if (digests_len == 0) { orc_hash_or_noop(leaves, leaf_len, out); return; }
    size_t half = digests_len / 2;
    gl_t* left_buf = digests;                 /* half-1 entries, then left child digest */
    gl_t* left_mem = digests + 4 * (half - 1);
    gl_t* right_mem = digests + 4 * half;
    gl_t* right_buf = digests + 4 * (half + 1);
    gl_t l[4], r[4];
    fill_subtree(left_buf, half - 1, leaves, n_leaves / 2, leaf_len, l);
    fill_subtree(right_buf, half - 1, leaves + (n_leaves / 2) * leaf_len, n_leaves / 2, leaf_len, r);
    memcpy(left_mem, l, sizeof l);
    memcpy(right_mem, r, sizeof r);
    orc_two_to_one(l, r, out);
}
/* returns 0 ok, -1 bad args (plonky2 asserts: power-of-two leaves, cap_height <= log2 leaves). */
int orc_merkle_new(const uint64_t* leaves, uint64_t n_leaves, uint32_t leaf_len, uint32_t cap_height,
                   uint64_t* digests, uint64_t* cap) {
    if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return -1;
    unsigned lg = 0;
    while (((uint64_t)1 << lg) < n_leaves) lg++;
    if (cap_height > lg) return -1;
    size_t n_cap = (size_t)1 << cap_height;
    size_t sub_leaves = n_leaves >> cap_height;
    size_t sub_digests = 2 * (sub_leaves - 1);
    for (size_t s = 0; s < n_cap; s++)
        fill_subtree(digests + 4 * s * sub_digests, sub_digests, leaves + s * sub_leaves * leaf_len,
                     sub_leaves, leaf_len, cap + 4 * s);
    return 0;
}
/* `MerkleTree::prove`: log2(N) - cap_height siblings, bottom-up (A11). */
void orc_merkle_prove(const uint64_t* digests, uint64_t n_leaves, uint32_t cap_height,
                      uint64_t leaf_index, uint64_t* siblings) {
    unsigned lg = 0;
    while (((uint64_t)1 << lg) < n_leaves) lg++;
    unsigned num_layers = lg - cap_height;
    size_t sub_digests = ((size_t)1 << (num_layers + 1)) - 2;
    size_t subtree = leaf_index >> num_layers;
    const uint64_t* tree = digests + 4 * subtree * sub_digests;
    size_t pair_index = leaf_index & (((size_t)1 << num_layers) - 1);
    for (unsigned i = 0; i < num_layers; i++) {
        size_t parity = pair_index & 1;
        pair_index >>= 1;
        size_t siblings_index = (pair_index << (i + 1)) + ((size_t)1 << i) - 1;
        size_t sibling_index = 2 * siblings_index + (1 - parity);
        memcpy(siblings + 4 * i, tree + 4 * sibling_index, 4 * sizeof(uint64_t));
    }
}
/* merkle_proofs.rs `verify_merkle_proof_to_cap`; returns 1 if the path leads to cap[index]. */
int orc_merkle_verify(const uint64_t* leaf, uint32_t leaf_len, uint64_t leaf_index,
                      const uint64_t* siblings, uint32_t n_siblings, const uint64_t* cap) {
    gl_t cur[4], nxt[4];
    orc_hash_or_noop(leaf, leaf_len, cur);
    uint64_t index = leaf_index;
    for (uint32_t i = 0; i < n_siblings; i++) {
        if (index & 1) orc_two_to_one(siblings + 4 * i, cur, nxt);
        else orc_two_to_one(cur, siblings + 4 * i, nxt);
        memcpy(cur, nxt, sizeof cur);
        index >>= 1;
    }
    return memcmp(cur, cap + 4 * index, sizeof cur) == 0;
}

/* ------------------------------------------------------------------ PolynomialBatch (A1, A3, A5)
 * fri/oracle.rs `from_values` (is_coeffs = 0) / `from_coeffs` (is_coeffs = 1).
 *   in        k columns x n, column-major
 *   salt      NULL, or 4 columns x N (natural LDE order, column-major) = the F::rand_vec(N) columns
 *             plonky2 chains after the polynomials when `blinding`
 *   coeffs    k x n column-major               (PolynomialBatch.polynomials)
 *   leaves    N rows x (k + salt) row-major    (merkle_tree.leaves; row j = LDE row bitrev(j))
 *   digests   4 * 2(N - 2^h), cap 4 * 2^h
 */
int orc_commit(const uint64_t* in, int is_coeffs, uint32_t n_log, uint32_t k, uint32_t rate_bits,
               uint32_t cap_height, const uint64_t* salt, uint64_t* coeffs, uint64_t* leaves,
               uint64_t* digests, uint64_t* cap) {
    size_t n = (size_t)1 << n_log, N = n << rate_bits;
    unsigned N_log = n_log + rate_bits;
    if (cap_height > N_log) return -1;
    size_t row = k + (salt ? 4 : 0);
    uint64_t* col = (uint64_t*)malloc(N * sizeof(uint64_t));
    if (!col) return -2;
    for (uint32_t c = 0; c < k; c++) {
        uint64_t* cf = coeffs + (size_t)c * n;
        for (size_t i = 0; i < n; i++) cf[i] = gl_canon(in[(size_t)c * n + i]);
        if (!is_coeffs) orc_ifft(cf, n_log);
        orc_coset_lde(cf, n_log, rate_bits, col);
        /* transpose + reverse_index_bits_in_place(leaves) */
        for (size_t j = 0; j < N; j++) leaves[j * row + c] = col[bitrev64(j, N_log)];
    }
    if (salt)
        for (uint32_t s = 0; s < 4; s++)
            for (size_t j = 0; j < N; j++)
                leaves[j * row + k + s] = gl_canon(salt[(size_t)s * N + bitrev64(j, N_log)]);
    free(col);
    return orc_merkle_new(leaves, N, (uint32_t)row, cap_height, digests, cap);
}

/* ------------------------------------------------------------------ openings (row N3)
 * plonky2/src/plonk/proof.rs `OpeningSet::new` -> `eval_commitment`: p.to_extension().eval(z) for every polynomial of a
 * batch, z in QuadraticExtension<GoldilocksField> = F_p[X]/(X^2 - W), W = 7 (field/src/extension/quadratic.rs,
 * goldilocks_extensions.rs).  Horner from the top coefficient.  PARITY UNPINNED (no fixture in the reference). */
void orc_eval_ext2(const uint64_t* coeffs, uint64_t n, const uint64_t zeta[2], uint64_t out[2]) {
    gl_t za = gl_canon(zeta[0]), zb = gl_canon(zeta[1]);
    gl_t a = 0, b = 0;
    for (uint64_t i = n; i-- > 0;) {
        /* (a + bX)(za + zbX) = (a za + 7 b zb) + (a zb + b za) X */
        gl_t na = gl_add(gl_mul(a, za), gl_mul(7, gl_mul(b, zb)));
        gl_t nb = gl_add(gl_mul(a, zb), gl_mul(b, za));
        a = gl_add(na, gl_canon(coeffs[i]));
        b = nb;
    }
    out[0] = a; out[1] = b;
}
