// C++ host-side mirror of plonky2's opening proof above the C ABI (rows N2 + N3 of SURVEY.md section 8f), same names and
// argument meaning as plonky2 @ f99ed9c (pinned by /root/reference/Cargo.toml:12):
//     Challenger                                   plonky2/src/iop/challenger.rs
//     FriConfig / FriParams / ConstantArityBits    plonky2/src/fri/mod.rs, fri/reduction_strategies.rs
//     FriInstanceInfo / FriBatchInfo / FriPolynomialInfo                      plonky2/src/fri/structure.rs
//     PolynomialBatch::prove_openings              plonky2/src/fri/oracle.rs
//     fri_proof, fri_committed_trees, fri_proof_of_work, fri_prover_query_rounds   plonky2/src/fri/prover.rs
// The transcript is sequenced here; every polynomial step is a b200zkp_fri_* call on data that stays in HBM.
// Header-only; link libb200zkp.so.  The Python twin (intmax_zkp_core_b200/fri.py) is what the parity tests drive.
#pragma once
#include <utility>

#include "plonky2_api.hpp"

namespace plonky2 {

using Ext = std::array<F, 2>;   // a + b X, X^2 = 7
constexpr size_t SPONGE_RATE = 8, SPONGE_WIDTH = 12;

class Challenger {
  public:
    explicit Challenger(const Context& c) : ctx_(&c) {}
    void observe_element(F e) {
        output_buffer_.clear();
        input_buffer_.push_back(e >= GOLDILOCKS_ORDER ? e - GOLDILOCKS_ORDER : e);
        if (input_buffer_.size() == SPONGE_RATE) duplexing();
    }
    template <class It> void observe_elements(It first, It last) { for (; first != last; ++first) observe_element(*first); }
    void observe_hash(const HashOut& h) { observe_elements(h.elements.begin(), h.elements.end()); }
    void observe_cap(const MerkleCap& cap) { for (const HashOut& h : cap) observe_hash(h); }
    void observe_extension_element(const Ext& e) { observe_elements(e.begin(), e.end()); }
    F get_challenge() {
        if (!input_buffer_.empty() || output_buffer_.empty()) duplexing();
        F v = output_buffer_.back();
        output_buffer_.pop_back();
        return v;
    }
    HashOut get_hash() { HashOut h; for (F& e : h.elements) e = get_challenge(); return h; }
    Ext get_extension_challenge() { Ext e; e[0] = get_challenge(); e[1] = get_challenge(); return e; }

  private:
    void duplexing() {
        for (size_t i = 0; i < input_buffer_.size(); i++) sponge_state_[i] = input_buffer_[i];   // overwrite mode
        input_buffer_.clear();
        std::array<F, SPONGE_WIDTH> out{};
        ctx_->check(b200zkp_poseidon_permute(ctx_->raw(), sponge_state_.data(), 1, out.data()));
        sponge_state_ = out;
        output_buffer_.assign(sponge_state_.begin(), sponge_state_.begin() + SPONGE_RATE);
    }
    const Context* ctx_;
    std::array<F, SPONGE_WIDTH> sponge_state_{};
    std::vector<F> input_buffer_, output_buffer_;
};

struct FriConfig {
    size_t rate_bits = 3, cap_height = 4, proof_of_work_bits = 16;
    size_t arity_bits = 4, final_poly_bits = 5;   // FriReductionStrategy::ConstantArityBits(4, 5)
    size_t num_query_rounds = 28;                 // = CircuitConfig::standard_recursion_config().fri_config
};

struct FriParams {
    FriConfig config;
    bool hiding = false;
    size_t degree_bits = 0;
    std::vector<size_t> reduction_arity_bits;
    size_t lde_bits() const { return degree_bits + config.rate_bits; }
    // FriConfig::fri_params with ConstantArityBits::reduction_arity_bits
    static FriParams from_config(const FriConfig& cfg, size_t degree_bits, bool hiding = false) {
        FriParams p;
        p.config = cfg; p.hiding = hiding; p.degree_bits = degree_bits;
        size_t d = degree_bits;
        while (d > cfg.final_poly_bits && d + cfg.rate_bits >= cfg.cap_height + cfg.arity_bits) {
            p.reduction_arity_bits.push_back(cfg.arity_bits);
            if (d < cfg.arity_bits) throw std::invalid_argument("degree_bits < arity_bits");
            d -= cfg.arity_bits;
        }
        return p;
    }
};

struct FriPolynomialInfo { size_t oracle_index, polynomial_index; };
struct FriBatchInfo { Ext point; std::vector<FriPolynomialInfo> polynomials; };
struct FriInstanceInfo { std::vector<FriBatchInfo> batches; };

struct FriQueryStep { std::vector<Ext> evals; MerkleProof merkle_proof; };
struct FriQueryRound {
    std::vector<std::pair<std::vector<F>, MerkleProof>> initial_trees_proof;   // FriInitialTreeProof::evals_proofs
    std::vector<FriQueryStep> steps;
};
struct FriProof {
    std::vector<MerkleCap> commit_phase_merkle_caps;
    std::vector<FriQueryRound> query_round_proofs;
    std::vector<Ext> final_poly;
    F pow_witness = 0;
};

// PolynomialBatch::prove_openings(instance, oracles, challenger, fri_params, timing).  mul_by_x: the pinned 2022 revision
// multiplies final_poly by X (plonky2 PR 436); later revisions pad the quotient instead.
inline FriProof prove_openings(const FriInstanceInfo& instance, const std::vector<const PolynomialBatch*>& oracles,
                               Challenger& challenger, const FriParams& fri_params, bool mul_by_x) {
    if (oracles.empty()) throw std::invalid_argument("no oracle");
    const Context& c = oracles[0]->context();
    const FriConfig& cfg = fri_params.config;
    for (const PolynomialBatch* o : oracles)
        if (o->degree_log != fri_params.degree_bits || o->rate_bits != cfg.rate_bits)
            throw std::invalid_argument("oracle shape does not match the FRI parameters");
    const Ext alpha = challenger.get_extension_challenge();

    std::vector<b200zkp_batch*> handles;
    for (const PolynomialBatch* o : oracles) handles.push_back(o->raw());
    std::vector<uint64_t> points;
    std::vector<uint32_t> counts, poly_oracle, poly_index;
    for (const FriBatchInfo& b : instance.batches) {
        points.push_back(b.point[0]); points.push_back(b.point[1]);
        counts.push_back((uint32_t)b.polynomials.size());
        for (const FriPolynomialInfo& p : b.polynomials) { poly_oracle.push_back((uint32_t)p.oracle_index); poly_index.push_back((uint32_t)p.polynomial_index); }
    }
    b200zkp_fri* raw = nullptr;
    c.check(b200zkp_fri_begin(c.raw(), handles.data(), (uint32_t)handles.size(), (uint32_t)counts.size(), points.data(), counts.data(),
                              poly_oracle.data(), poly_index.data(), alpha.data(), mul_by_x ? B200ZKP_FRI_MUL_BY_X : 0, &raw));
    std::shared_ptr<b200zkp_fri> fri(raw, b200zkp_fri_free);

    FriProof proof;
    // fri_committed_trees
    for (size_t arity_bits : fri_params.reduction_arity_bits) {
        MerkleCap cap(size_t(1) << cfg.cap_height);
        c.check(b200zkp_fri_commit_layer(fri.get(), (uint32_t)arity_bits, (uint32_t)cfg.cap_height, cap[0].elements.data()));
        challenger.observe_cap(cap);
        proof.commit_phase_merkle_caps.push_back(cap);
        const Ext beta = challenger.get_extension_challenge();
        c.check(b200zkp_fri_fold(fri.get(), beta.data()));
    }
    uint32_t shape[4];
    c.check(b200zkp_fri_shape(fri.get(), shape));
    proof.final_poly.resize((size_t(1) << shape[2]) >> cfg.rate_bits);
    c.check(b200zkp_fri_final_poly(fri.get(), proof.final_poly[0].data()));
    for (const Ext& e : proof.final_poly) challenger.observe_extension_element(e);
    // fri_proof_of_work: hash_no_pad(current_hash || w).elements[0] with proof_of_work_bits leading zeros, smallest w
    const HashOut current_hash = challenger.get_hash();
    std::array<F, SPONGE_WIDTH> state{};
    std::copy(current_hash.elements.begin(), current_hash.elements.end(), state.begin());
    c.check(b200zkp_pow_grind(c.raw(), state.data(), 4, 0, (uint32_t)cfg.proof_of_work_bits, 0, &proof.pow_witness));
    // fri_prover_query_rounds: the indices do not depend on the answers, so the gathers are batched per tree
    const size_t nq = cfg.num_query_rounds;
    const uint64_t lde_size = uint64_t(1) << fri_params.lde_bits();
    std::vector<uint64_t> idx(nq);
    for (uint64_t& x : idx) x = challenger.get_challenge() % lde_size;
    proof.query_round_proofs.resize(nq);
    for (const PolynomialBatch* o : oracles) {
        size_t row = o->num_polys + o->salt_size, depth = o->degree_log + o->rate_bits - o->cap_height;
        std::vector<F> rows(nq * row);
        std::vector<HashOut> sib(nq * depth);
        c.check(b200zkp_batch_rows(o->raw(), idx.data(), nq, rows.data(), depth ? sib[0].elements.data() : nullptr));
        for (size_t r = 0; r < nq; r++) {
            MerkleProof p;
            p.siblings.assign(sib.begin() + r * depth, sib.begin() + (r + 1) * depth);
            proof.query_round_proofs[r].initial_trees_proof.emplace_back(std::vector<F>(rows.begin() + r * row, rows.begin() + (r + 1) * row), p);
        }
    }
    size_t bits = fri_params.lde_bits();
    for (size_t layer = 0; layer < fri_params.reduction_arity_bits.size(); layer++) {
        size_t ab = fri_params.reduction_arity_bits[layer], arity = size_t(1) << ab;
        for (uint64_t& x : idx) x >>= ab;
        bits -= ab;
        size_t depth = bits - cfg.cap_height;
        std::vector<Ext> evals(nq * arity);
        std::vector<HashOut> sib(nq * depth);
        c.check(b200zkp_fri_query(fri.get(), (uint32_t)layer, idx.data(), nq, evals[0].data(), depth ? sib[0].elements.data() : nullptr));
        for (size_t r = 0; r < nq; r++) {
            FriQueryStep st;
            st.evals.assign(evals.begin() + r * arity, evals.begin() + (r + 1) * arity);
            st.merkle_proof.siblings.assign(sib.begin() + r * depth, sib.begin() + (r + 1) * depth);
            proof.query_round_proofs[r].steps.push_back(std::move(st));
        }
    }
    return proof;
}

}  // namespace plonky2
