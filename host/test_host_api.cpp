// Exercises host/plonky2_api.hpp on cuda:0: the reference's own Poseidon fixture, SURVEY.md App. C commitment,
// Merkle proofs.  Built and run by tests/test_host_cpp.py (gpu); exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include "fri_api.hpp"
#include "prover_api.hpp"

using namespace plonky2;

static uint64_t splitmix64(uint64_t j) {
    uint64_t z = (j + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main() {
    Context ctx(0);
    // /root/reference/src/transaction/circuits/mod.rs:211-218
    HashOut z{};
    HashOut h = PoseidonHash::two_to_one(ctx, z, z);
    CHECK((h.elements == std::array<F, 4>{4330397376401421145ull, 14124799381142128323ull, 8742572140681234676ull, 14345658006221440202ull}));
    // n = 8, k = 9, rate_bits = 3, cap_height = 2
    const size_t n = 8, k = 9;
    std::vector<std::vector<F>> values(k, std::vector<F>(n));
    for (size_t c = 0; c < k; c++)
        for (size_t i = 0; i < n; i++) { uint64_t v = splitmix64(c * n + i); values[c][i] = v >= GOLDILOCKS_ORDER ? v - GOLDILOCKS_ORDER : v; }
    PolynomialBatch b = PolynomialBatch::from_values(ctx, values, 3, false, 2);
    CHECK(b.cap.size() == 4);
    CHECK((b.cap[0].elements == std::array<F, 4>{9531016979423918488ull, 17086599980695735262ull, 12854109491395286945ull, 1436292215001049984ull}));
    CHECK((b.cap[3].elements == std::array<F, 4>{7263687849510528849ull, 3413957559828268007ull, 13767290799342735489ull, 13701751426046424270ull}));
    CHECK(b.polynomials()[0][0] == 8511423253799370256ull);
    for (uint64_t j : {0ull, 1ull, 37ull, 63ull}) {
        auto opened = b.open(j);
        CHECK(opened.first.size() == k && opened.second.siblings.size() == 4);
        verify_merkle_proof_to_cap(ctx, opened.first, j, b.cap, opened.second);
        bool threw = false;
        try { verify_merkle_proof_to_cap(ctx, opened.first, j ^ 1, b.cap, opened.second); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
    }
    CHECK(b.get_lde_values(0)[0] == 1132379675625856675ull);       // leaf[0][0]
    // MerkleTree::new over the same leaves gives the same cap; plonky2's asserts surface as invalid_argument
    std::vector<std::vector<F>> leaves;
    for (uint64_t j = 0; j < 64; j++) leaves.push_back(b.open(j).first);
    MerkleTree t = MerkleTree::new_(ctx, leaves, 2);
    CHECK(t.cap == b.cap);
    CHECK(t.prove(5).siblings == b.open(5).second.siblings);
    bool threw = false;
    try { MerkleTree::new_(ctx, leaves, 7); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { leaves.pop_back(); MerkleTree::new_(ctx, leaves, 2); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    // opening proof (rows N2 + N3): two oracles of 2^6 coefficients, a small FRI configuration; the digest of the whole
    // proof is compared with the Python mirror (itself bit-exact against oracle/fri_ref.py) by tests/test_host_cpp.py
    {
        const size_t fn = 64;
        auto synth = [&](size_t kk, uint64_t seed) {
            std::vector<std::vector<F>> v(kk, std::vector<F>(fn));
            for (size_t c = 0; c < kk; c++)
                for (size_t i = 0; i < fn; i++) { uint64_t x = splitmix64((seed << 48) + c * fn + i); v[c][i] = x >= GOLDILOCKS_ORDER ? x - GOLDILOCKS_ORDER : x; }
            return v;
        };
        FriConfig cfg;
        cfg.rate_bits = 2; cfg.cap_height = 1; cfg.proof_of_work_bits = 7; cfg.arity_bits = 2; cfg.final_poly_bits = 2; cfg.num_query_rounds = 5;
        FriParams params = FriParams::from_config(cfg, 6);
        CHECK(params.reduction_arity_bits.size() == 2);
        PolynomialBatch o0 = PolynomialBatch::from_coeffs(ctx, synth(3, 5), cfg.rate_bits, false, cfg.cap_height);
        PolynomialBatch o1 = PolynomialBatch::from_coeffs(ctx, synth(2, 6), cfg.rate_bits, false, cfg.cap_height);
        Challenger ch(ctx);
        ch.observe_cap(o0.cap);
        ch.observe_cap(o1.cap);
        FriInstanceInfo inst;
        inst.batches.push_back({Ext{123456789ull, 987654321ull}, {{0, 0}, {0, 1}, {0, 2}, {1, 0}, {1, 1}}});
        inst.batches.push_back({Ext{555ull, 777ull}, {{1, 0}, {1, 1}}});
        FriProof proof = prove_openings(inst, {&o0, &o1}, ch, params, /*mul_by_x=*/true);
        CHECK(proof.commit_phase_merkle_caps.size() == 2 && proof.final_poly.size() == 4 && proof.query_round_proofs.size() == 5);
        uint64_t d = 0xcbf29ce484222325ull;
        auto mix = [&](uint64_t w) { d = (d ^ w) * 0x100000001b3ull; };
        for (auto& cap : proof.commit_phase_merkle_caps) for (auto& hh : cap) for (F e : hh.elements) mix(e);
        for (auto& e : proof.final_poly) { mix(e[0]); mix(e[1]); }
        mix(proof.pow_witness);
        for (auto& r : proof.query_round_proofs) {
            for (auto& ip : r.initial_trees_proof) { for (F e : ip.first) mix(e); for (auto& sb : ip.second.siblings) for (F e : sb.elements) mix(e); }
            for (auto& st : r.steps) { for (auto& e : st.evals) { mix(e[0]); mix(e[1]); } for (auto& sb : st.merkle_proof.siblings) for (F e : sb.elements) mix(e); }
        }
        mix(ch.get_challenge());
        std::printf("fri digest %016llx pow %llu\n", (unsigned long long)d, (unsigned long long)proof.pow_witness);
        auto ev = o1.eval_ext2(Ext{555ull, 777ull});
        CHECK(ev.size() == 2);
    }
    {   // row N1a: identity permutation (sigma = k_j * w^i) makes every quotient 1, so Z and all partial products are 1;
        // swapping two equal wires keeps the argument closed, a wrong sigma opens it
        const size_t R = 16, nn = 32, deg = 8;
        auto k_is = get_unique_coset_shifts(R);
        CHECK(k_is[2] == 49 && num_partial_products(80, 8) == 9);
        auto mulmod = [](F a, F b) { return (F)((unsigned __int128)a * b % GOLDILOCKS_ORDER); };
        F w32 = 1753635133440165772ull;                      // generator of the 2^32 subgroup
        for (int i = 5; i < 32; i++) w32 = mulmod(w32, w32);  // w_32
        std::vector<std::vector<F>> wires(R, std::vector<F>(nn)), sigmas(R, std::vector<F>(nn));
        for (size_t j = 0; j < R; j++) {
            F x = 1;
            for (size_t i = 0; i < nn; i++) { wires[j][i] = splitmix64(j * nn + i) % GOLDILOCKS_ORDER; sigmas[j][i] = mulmod(k_is[j], x); x = mulmod(x, w32); }
        }
        auto cols = zs_partial_products(ctx, wires, sigmas, k_is, {11, 22}, {33, 44}, deg);
        CHECK(cols.size() == 4 && cols[0].size() == nn);
        for (auto& col : cols) for (F v : col) CHECK(v == 1);
        wires[3][7] = wires[9][20];                          // copy constraint (3, 7) <-> (9, 20)
        std::swap(sigmas[3][7], sigmas[9][20]);
        cols = zs_partial_products(ctx, wires, sigmas, k_is, {11, 22}, {33, 44}, deg);
        CHECK(cols[0][0] == 1 && cols[0][8] != 1 && cols[0][21] == 1 && cols[1][31] == 1);
        auto per = all_wires_permutation_partial_products(ctx, wires, sigmas, k_is, {11, 22}, {33, 44}, deg);
        CHECK(per.size() == 2 && per[1].size() == 2 && per[1][1] == cols[1] && per[0][0] == cols[2]);
    }
    if (b200zkp_device_count() >= 2) {
        // one process, every GPU of the box (b200zkp_comm_init_all): same cap, same openings as the single-GPU commitment
        int G = 2;
        while (G * 2 <= b200zkp_device_count() && G * 2 <= 8) G *= 2;
        std::vector<std::unique_ptr<Context>> owned;
        std::vector<const Context*> cs;
        for (int g = 0; g < G; g++) { owned.emplace_back(new Context(g)); cs.push_back(owned.back().get()); }
        const size_t sn = 1 << 9, sk = 21;
        std::vector<std::vector<F>> sv(sk, std::vector<F>(sn));
        for (size_t c = 0; c < sk; c++)
            for (size_t i = 0; i < sn; i++) { uint64_t v = splitmix64((7ull << 48) + c * sn + i); sv[c][i] = v >= GOLDILOCKS_ORDER ? v - GOLDILOCKS_ORDER : v; }
        PolynomialBatch one = PolynomialBatch::from_values(ctx, sv, 3, false, 4);
        ShardedPolynomialBatch many = ShardedPolynomialBatch::from_values(cs, sv, 3, 4);
        CHECK(many.cap == one.cap);
        for (uint64_t j : {0ull, 1ull, 2047ull, 2048ull, 4095ull}) {
            auto a = one.open(j), b2 = many.open(j);
            CHECK(a.first == b2.first && a.second.siblings == b2.second.siblings);
        }
        bool bad = false;
        try { ShardedPolynomialBatch::from_values(cs, sv, 0, 4); } catch (const std::invalid_argument&) { bad = true; }
        CHECK(bad);
        std::printf("multi-GPU C++ API ok on %d devices\n", G);
    }
    std::printf("host C++ API ok\n");
    return 0;
}
