// C++ host-side mirror of the prover step between the wires commitment and the Z commitment (row N1a), above the C ABI.
//
//     all_wires_permutation_partial_products / wires_permutation_partial_products_and_zs     plonky2/src/plonk/prover.rs
//     get_unique_coset_shifts (CommonCircuitData::k_is), num_partial_products               plonky2/src/plonk/plonk_common.rs
// (plonky2 @ f99ed9c, pinned by /root/reference/Cargo.toml:12; reached from every prove(), e.g.
// /root/reference/src/transaction/circuits/mod.rs:453).  Same names and argument meaning; header-only.
#pragma once
#include "plonky2_api.hpp"

namespace plonky2 {

inline std::vector<F> get_unique_coset_shifts(size_t num_shifts) {
    std::vector<F> out(num_shifts);
    unsigned __int128 x = 1;
    for (size_t j = 0; j < num_shifts; j++) { out[j] = (F)x; x = x * 7 % GOLDILOCKS_ORDER; }
    return out;
}

inline size_t num_partial_products(size_t num_routed_wires, size_t quotient_degree_factor) {
    return (num_routed_wires + quotient_degree_factor - 1) / quotient_degree_factor - 1;
}

// The `Z + partial products` batch of prove(), in committed order: Z of every challenge, then the partial products
// challenge by challenge.  wires, sigmas: one vector of n values per routed wire.
inline std::vector<std::vector<F>> zs_partial_products(const Context& c, const std::vector<std::vector<F>>& wires,
                                                       const std::vector<std::vector<F>>& sigmas, const std::vector<F>& k_is,
                                                       const std::vector<F>& betas, const std::vector<F>& gammas,
                                                       size_t quotient_degree_factor) {
    if (wires.empty() || wires.size() != sigmas.size() || k_is.size() != wires.size() || betas.empty() || betas.size() != gammas.size())
        throw std::invalid_argument("wires, sigmas and k_is need one entry per routed wire; betas and gammas one per challenge");
    const size_t R = wires.size(), n = wires[0].size(), n_log = log2_strict(n);
    std::vector<F> w(R * n), s(R * n);
    for (size_t j = 0; j < R; j++) {
        if (wires[j].size() != n || sigmas[j].size() != n) throw std::invalid_argument("columns of different lengths");
        std::copy(wires[j].begin(), wires[j].end(), w.begin() + j * n);
        std::copy(sigmas[j].begin(), sigmas[j].end(), s.begin() + j * n);
    }
    const size_t chunks = (R + quotient_degree_factor - 1) / quotient_degree_factor, cols = betas.size() * chunks;
    std::vector<F> out(cols * n);
    c.check(b200zkp_partial_products_and_zs(c.raw(), w.data(), s.data(), (uint32_t)n_log, (uint32_t)R, (uint32_t)quotient_degree_factor,
                                            k_is.data(), betas.data(), gammas.data(), (uint32_t)betas.size(), out.data()));
    std::vector<std::vector<F>> res(cols);
    for (size_t q = 0; q < cols; q++) res[q].assign(out.begin() + q * n, out.begin() + (q + 1) * n);
    return res;
}

// plonky2's per-challenge result: [pp_0 .. pp_{num_prods - 1}, Z] for every challenge
inline std::vector<std::vector<std::vector<F>>> all_wires_permutation_partial_products(
    const Context& c, const std::vector<std::vector<F>>& wires, const std::vector<std::vector<F>>& sigmas, const std::vector<F>& k_is,
    const std::vector<F>& betas, const std::vector<F>& gammas, size_t quotient_degree_factor) {
    auto cols = zs_partial_products(c, wires, sigmas, k_is, betas, gammas, quotient_degree_factor);
    const size_t C = betas.size(), num_prods = cols.size() / C - 1;
    std::vector<std::vector<std::vector<F>>> res(C);
    for (size_t ch = 0; ch < C; ch++) {
        for (size_t l = 0; l < num_prods; l++) res[ch].push_back(cols[C + ch * num_prods + l]);
        res[ch].push_back(cols[ch]);
    }
    return res;
}

// PolynomialValues::coset_ifft(shift) over k vectors of n values (plonky2_field polynomial/mod.rs), row N1c
inline std::vector<std::vector<F>> coset_ifft_batch(const Context& c, const std::vector<std::vector<F>>& values, F shift) {
    if (values.empty()) return {};
    const size_t k = values.size(), n = values[0].size(), n_log = log2_strict(n);
    std::vector<F> flat(k * n);
    for (size_t j = 0; j < k; j++) {
        if (values[j].size() != n) throw std::invalid_argument("columns of different lengths");
        std::copy(values[j].begin(), values[j].end(), flat.begin() + j * n);
    }
    c.check(b200zkp_coset_intt(c.raw(), flat.data(), (uint32_t)n_log, (uint32_t)k, shift));
    std::vector<std::vector<F>> out(k);
    for (size_t j = 0; j < k; j++) out[j].assign(flat.begin() + j * n, flat.begin() + (j + 1) * n);
    return out;
}

// tail of compute_quotient_polys + the chunking in prove(): one vector of n << q quotient values per challenge (coset 7 <w>,
// natural order) -> the degree-n coefficient chunks PolynomialBatch::from_coeffs takes
// quotient_degree_factor (CommonCircuitData; need not be a power of two): prove() runs
// `quotient_poly.trim_to_len(quotient_degree_factor * degree)`, which panics when a dropped coefficient is non-zero (the
// witness does not satisfy the circuit), and commits exactly quotient_degree_factor chunks per challenge
inline std::vector<std::vector<F>> quotient_poly_chunks(const Context& c, const std::vector<std::vector<F>>& quotient_values,
                                                        size_t degree, size_t quotient_degree_factor) {
    auto coeffs = coset_ifft_batch(c, quotient_values, 7);
    std::vector<std::vector<F>> chunks;
    for (auto& p : coeffs) {
        if (degree == 0 || p.size() % degree) throw std::invalid_argument("quotient length must be a multiple of the degree");
        if (quotient_degree_factor == 0 || quotient_degree_factor * degree > p.size())
            throw std::invalid_argument("quotient_degree_factor * degree exceeds the quotient length");
        for (size_t i = quotient_degree_factor * degree; i < p.size(); i++)
            if (p[i] != 0) throw std::runtime_error("Quotient has failed, the vanishing polynomial is not divisible by Z_H");
        for (size_t o = 0; o < quotient_degree_factor * degree; o += degree) chunks.emplace_back(p.begin() + o, p.begin() + o + degree);
    }
    return chunks;
}

}  // namespace plonky2
