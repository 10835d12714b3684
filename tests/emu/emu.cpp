// TEST HARNESS (not product): compiles the CUDA kernel bodies of intmax_zkp_core_b200/csrc with g++ under
// B200ZKP_HOST_EMU and steps the "threads" of each CTA in a loop, so the index logic of the NTT passes, the
// Poseidon permutation arithmetic and the digest layout can be checked on a machine without a GPU.
// The product library never defines B200ZKP_HOST_EMU and has no CPU path.
#define B200ZKP_HOST_EMU 1
#define __restrict__
#include <cstring>
#include <vector>

#include "../../intmax_zkp_core_b200/csrc/merkle_kernels.cuh"
#include "../../intmax_zkp_core_b200/csrc/host_plan.hpp"

using gl::u32;
using gl::u64;

static void run_pass(const ntt::PassParams& p, u32 B) {
    u64 T = ntt::TILE_ELEMS >> B;
    u64 total_batches = (u64)p.ncols << (p.n_log - B);
    u64 blocks = (total_batches + T - 1) / T;
    for (u64 blk = 0; blk < blocks; blk++) {
        switch (B) {
            case 1: ntt::pass_body<1>(p, (u32)blk); break;
            case 2: ntt::pass_body<2>(p, (u32)blk); break;
            case 3: ntt::pass_body<3>(p, (u32)blk); break;
            case 4: ntt::pass_body<4>(p, (u32)blk); break;
            case 5: ntt::pass_body<5>(p, (u32)blk); break;
            case 6: ntt::pass_body<6>(p, (u32)blk); break;
            case 7: ntt::pass_body<7>(p, (u32)blk); break;
            case 8: ntt::pass_body<8>(p, (u32)blk); break;
        }
    }
}

struct Tables {
    std::vector<u64> w[2][9];
    Tables() { for (int d = 0; d < 2; d++) for (u32 B = 1; B <= 8; B++) w[d][B] = hostgl::small_root_table(B, d); }
};
static Tables g_tab;

static void transform(const u64* in, u64 in_stride, u64* out, u64 out_stride, u64* scratch, u32 n_log, u32 ncols,
                      int dir, bool bitrev_out, bool has_scale, u64 scale_base, u64 out_scale) {
    if (n_log == 0) { for (u32 c = 0; c < ncols; c++) out[c * out_stride] = gl::canon(in[c * in_stride]); return; }
    u64* wt[9];
    for (u32 B = 0; B <= 8; B++) wt[B] = B ? g_tab.w[dir][B].data() : nullptr;
    std::vector<u64> tlo, thi, slo, shi;
    u64 w = hostgl::root(n_log);
    if (dir) w = hostgl::inv(w);
    ntt::TwoLevelPtr tw{nullptr, nullptr, 0}, sc{nullptr, nullptr, 0};
    tw.lo_bits = hostgl::two_level_powers(w, n_log, &tlo, &thi); tw.lo = tlo.data(); tw.hi = thi.data();
    if (has_scale) { sc.lo_bits = hostgl::two_level_powers(scale_base, n_log, &slo, &shi); sc.lo = slo.data(); sc.hi = shi.data(); }
    ntt::Plan plan;
    ntt::make_plan(&plan, in, in_stride, out, out_stride, scratch, n_log, ncols, wt, tw, bitrev_out,
                   has_scale ? &sc : nullptr, out_scale, true);
    for (u32 pi = 0; pi < plan.n_passes; pi++) run_pass(plan.pass[pi], plan.bits[pi]);
}

extern "C" {
void emu_permute(u64* s) { u64 t[12]; memcpy(t, s, sizeof t); poseidon::permute(t); memcpy(s, t, sizeof t); }
void emu_sponge(const u64* leaf, u64 col_stride, u32 len, u32 noop_short, u64* out4) {
    u64 s[12];
    merkle::sponge_leaf(leaf, col_stride, len, noop_short, s);
    memcpy(out4, s, 32);
}
u64 emu_node_slot(u32 sub_log, u64 subtree, u32 layer, u64 m) {
    merkle::TreeShape t; t.sub_log = sub_log; t.sub_digests = 2 * (((u64)1 << sub_log) - 1);
    return merkle::node_slot(t, subtree, layer, m);
}
// values [k][n] -> coeffs [k][n]
void emu_intt(const u64* in, u64* out, u32 n_log, u32 k) {
    u64 n = (u64)1 << n_log;
    std::vector<u64> scratch((size_t)k * n);
    u64 n_inv = hostgl::inv(n % hostgl::P);
    transform(in, n, out, n, scratch.data(), n_log, k, 1, false, false, 0, n_log ? n_inv : 0);
}
void emu_ntt(const u64* in, u64* out, u32 n_log, u32 k) {
    u64 n = (u64)1 << n_log;
    std::vector<u64> scratch((size_t)k * n);
    transform(in, n, out, n, scratch.data(), n_log, k, 0, false, false, 0, 0);
}
// coeffs [k][n] -> lde [k][N] in leaf order
void emu_lde(const u64* coeffs, u64* lde, u32 n_log, u32 k, u32 rate_bits) {
    u64 n = (u64)1 << n_log, N = n << rate_bits;
    for (u32 b = 0; b < (1u << rate_bits); b++)
        transform(coeffs, n, lde + b * n, N, nullptr, n_log, k, 0, true, true,
                  hostgl::coset_shift_of_block(n_log, rate_bits, b), 0);
}
}
