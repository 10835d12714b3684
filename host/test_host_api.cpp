// Exercises host/plonky2_api.hpp on cuda:0: the reference's own Poseidon fixture, SURVEY.md App. C commitment,
// Merkle proofs.  Built and run by tests/test_host_cpp.py (gpu); exit code 0 = all checks passed.
#include <cstdio>
#include <cstdlib>
#include "plonky2_api.hpp"

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
    std::printf("host C++ API ok\n");
    return 0;
}
