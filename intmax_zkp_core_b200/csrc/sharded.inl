// One commitment partitioned over several GPUs behind the C ABI (include/b200zkp.h, "partitioned over several GPUs").
// Included at the end of b200zkp.cu (uses its static stage functions).  Product code: NCCL is the only library on this
// path and in the default (peer-memory) form it only moves the 512-byte cap and the handshakes of the set-up; the coefficient
// shards cross NVLink inside this repo's own first-pass kernel (ntc::ct_pull_kernel), every transform and hash is this
// repo's kernels.  The NCCL point-to-point form of the exchange is the fallback.
//
// Replaces nothing in plonky2 (the CPU prover is one process on one memory): it is north_star's multi-GPU form of
// PolynomialBatch::from_values, reached from the same prove() call sites
// (/root/reference/src/rollup/circuits/mod.rs:1247, src/transaction/circuits/mod.rs:453).
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <memory>
#include <thread>

#include "peer_kernels.cuh"

namespace {

// ---- libnccl.so.2, loaded on first use (the single-GPU ABI must load and run without NCCL)
struct NcclApi {
    void* handle = nullptr;
    std::string err;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    bool ok() const { return handle != nullptr && err.empty(); }
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_NOLOAD first: a host process that already mapped libnccl.so.2 (torch ships its own) must not get a second copy
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { const char* e = dlerror(); api.err = std::string("cannot load libnccl.so.2: ") + (e ? e : "?"); return; }
        api.handle = h;
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(h, name);
            if (!p && api.err.empty()) api.err = std::string("libnccl lacks ") + name;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    });
    return api;
}

}  // namespace

// Peer-memory exchange (the default inside one NVLink domain): every rank owns an "exchange window" that holds the coefficients
// of its own columns and that every other rank can read (peer access in one process, CUDA IPC between processes), plus the flag
// words of peer_kernels.cuh.  The all-gather of the coefficients then happens inside the first pass of the coset transforms
// (ntc::ct_pull_kernel): no NCCL call, no copy kernel, no staging of the shards between the exchange and the transform.
struct PeerRank {
    u64* win = nullptr;                 // this rank's window, [kp][n] of the largest shape created so far
    size_t win_b = 0;
    u32* flags = nullptr;               // peer::FLAG_WORDS words
    std::vector<const u64*> peer_win;   // [world]: rank s's window as seen from this rank's device (own rank: win)
    peer::PeerFlags peer_flags{};       // [world]: rank s's flags as seen from this rank's device
    std::vector<void*> imported_win, imported_flags;   // CUDA IPC mappings to close
};

// rendezvous of the host threads that drive the ranks of a one-process communicator (CUDA events order work between the
// devices of one process, but a wait must be issued after the record it refers to); abort() releases everyone when a rank fails
struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n = 1, arrived = 0;
    u64 gen = 0;
    bool broken = false;
    void reset(int n_) { std::lock_guard<std::mutex> lk(m); n = n_; arrived = 0; broken = false; }
    bool wait() {
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const u64 g0 = gen;
        if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return gen != g0 || broken; });
        return !broken;
    }
    void abort() { std::lock_guard<std::mutex> lk(m); broken = true; cv.notify_all(); }
};

struct b200zkp_comm {
    int world = 0;
    std::vector<int> rank;             // global rank of local rank i
    std::vector<b200zkp_ctx*> ctx;     // (not owned)
    std::vector<ncclComm_t> nc;
    std::vector<cudaStream_t> xstream; // side stream per local rank: NCCL point-to-point groups; in-place passes beside the gather pass
    std::vector<cudaStream_t> hstream; // hash stream per local rank: sponge absorption of a column chunk beside the transforms of the next
    std::vector<cudaEvent_t> ev;       // per local rank: [2 + world] events (coefficients ready, buffers free, one per group)
    uint32_t peers_per_group = 2;
    bool peer_ok = false;              // peer memory reaches every rank from every rank (decided collectively at init)
    bool peer_on = true;               // b200zkp_comm_set_peer_exchange / B200ZKP_PEER_EXCHANGE=0: use the NCCL exchange instead
    u32 epoch = 0;                     // commits issued on this communicator (the same number on every rank)
    u32 pull_groups = 4;               // device inputs: column groups of the gather / in-place pipeline (B200ZKP_PEER_GROUPS)
    u64 chunk_bytes = (u64)16 << 20;   // host inputs: bytes per upload chunk of the pipeline (B200ZKP_PEER_CHUNK_BYTES, for tests)
    std::vector<PeerRank> pr;          // per local rank
    // one process driving every rank: the ordering between devices is CUDA events (no spinning kernels), issued in step
    std::vector<cudaEvent_t> ready_ev; // [local rank * MAX_CHUNKS + chunk]
    std::vector<cudaEvent_t> done_ev;  // [local rank]
    bool done_valid = false;
    HostBarrier hb;
    bool one_process() const { return n_local() == world; }
    std::mutex mu;                     // serialises the collective entry points of this process
    std::string err;
    int n_local() const { return (int)rank.size(); }
    bool peer_exchange() const { return peer_ok && peer_on; }
};

struct ShardRank {
    u64 *coeffs_all = nullptr, *lde = nullptr, *digests = nullptr, *cap_local = nullptr, *cap = nullptr;
    u64 *stage = nullptr;     // host inputs: upload staging, [kp][n]
    u64 *gather = nullptr;    // NCCL exchange: the shards of every rank as received, [G][kp][n] (own shard at g * kp * n)
    u64 *sponge = nullptr;    // host inputs: the 12-word sponge state of every local leaf between column chunks, [12][N_local]
    size_t coeffs_b = 0, lde_b = 0, digests_b = 0, cap_local_b = 0, cap_b = 0, stage_b = 0, gather_b = 0, sponge_b = 0;
};

// Partition of the k columns over the G ranks: groups of L = max(8, G) columns (a multiple of the sponge rate, so a group can be
// hashed as soon as it is complete), w = L / G consecutive columns of every group per rank.  Column c belongs to rank
// (c % L) / w and is that rank's local column (c / L) * w + c % w; rank g's local column t is column (t / w) * L + g * w + t % w.
struct b200zkp_sharded {
    b200zkp_comm* comm = nullptr;
    u32 n_log = 0, k = 0, rate_bits = 0, cap_height = 0;
    u32 L = 0, w = 0, n_groups = 0;
    u32 kp = 0, bpr = 0, cpr = 0, cap_height_local = 0;      // kp = n_groups * w: local columns of a rank at most
    u64 N_local = 0;
    std::vector<ShardRank> r;
    u32 local_cols(u32 g) const {
        const u32 full = k / L, rem = k % L;
        return full * w + (rem > g * w ? std::min(w, rem - g * w) : 0u);
    }
    u32 global_col(u32 g, u32 t) const { return (t / w) * L + g * w + t % w; }
};

#define COMM_BAD(c, msg) do { (c)->err = (msg); return B200ZKP_ERR_BAD_ARG; } while (0)

static int nccl_fail(std::string* err, const char* what, ncclResult_t r) {
    NcclApi& a = nccl_api();
    *err = std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "nccl error");
    return B200ZKP_ERR_NCCL;
}
#define NCCL_TRY(errp, expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return nccl_fail((errp), #expr, r__); } while (0)

static int nccl_ready(std::string* err) {
    NcclApi& a = nccl_api();
    if (!a.ok()) { *err = a.err.empty() ? "NCCL unavailable" : a.err; return B200ZKP_ERR_NCCL; }
    return 0;
}

extern "C" int b200zkp_comm_unique_id(uint8_t id[B200ZKP_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == B200ZKP_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id) return B200ZKP_ERR_BAD_ARG;
    std::string err;
    if (nccl_ready(&err)) return B200ZKP_ERR_NCCL;
    ncclUniqueId u;
    if (nccl_api().GetUniqueId(&u) != ncclSuccess) return B200ZKP_ERR_NCCL;
    memcpy(id, &u, sizeof(u));
    return 0;
}

// runs f(local rank index) for every local rank: inline for one rank, one host thread per rank otherwise (each rank has its
// own device, ctx, NCCL communicator and streams; NCCL point-to-point groups of different ranks must be issued concurrently)
template <typename F>
static int for_each_rank(b200zkp_comm* c, F f) {
    const int n = c->n_local();
    if (n == 1) {
        int rc1 = f(0);
        if (rc1) c->err = c->ctx[0]->err;
        return rc1;
    }
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    th.reserve(n);
    for (int i = 0; i < n; i++) th.emplace_back([&, i] { rc[i] = f(i); });
    for (auto& t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i]) { c->err = "rank " + std::to_string(c->rank[i]) + ": " + c->ctx[i]->err; return rc[i]; }
    return 0;
}

// ---------------------------------------------------------------------------------------------- peer-memory exchange
static constexpr u64 PEER_WAIT_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// every rank of the communicator agrees on the AND of `ok` (one small NCCL all-reduce per local rank; also a barrier)
static int comm_all_ok(b200zkp_comm* c, const std::vector<int>& ok_local, bool* all) {
    std::vector<u32> res(c->n_local(), 0);
    int rc = for_each_rank(c, [&](int i) -> int {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        u32* d = nullptr;
        TRY(dev_alloc(ctx, 256, (void**)&d));
        u32 v = ok_local[i] ? 1u : 0u;
        int r = 0;
        if (cudaMemcpyAsync(d, &v, 4, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) r = B200ZKP_ERR_CUDA;
        if (!r) {
            ncclResult_t nr = nccl_api().AllReduce(d, d, 1, ncclUint32, ncclMin, c->nc[i], ctx->stream);
            if (nr != ncclSuccess) r = nccl_fail(&ctx->err, "ncclAllReduce", nr);
        }
        if (!r && (cudaMemcpyAsync(&res[i], d, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                   cudaStreamSynchronize(ctx->stream) != cudaSuccess)) r = B200ZKP_ERR_CUDA;
        if (r == B200ZKP_ERR_CUDA) { ctx->err = "CUDA error in the communicator handshake"; (void)cudaGetLastError(); }
        dev_release(ctx, d, 256);
        return r;
    });
    if (rc) return rc;
    *all = true;
    for (u32 v : res) *all = *all && v == 1u;
    return 0;
}

// own[i]: a whole cudaMalloc allocation of local rank i.  Fills views[i][s] = rank s's allocation as addressable from local
// rank i's device: the pointer itself inside one process (peer access), a CUDA IPC mapping between processes (recorded in
// imported[i]).  ok_local[i] is cleared when a mapping cannot be made; the caller decides collectively what to do then.
static int comm_exchange_ptrs(b200zkp_comm* c, const std::vector<void*>& own, std::vector<std::vector<void*>>* views,
                              std::vector<std::vector<void*>>* imported, std::vector<int>* ok_local) {
    const int nl = c->n_local(), W = c->world;
    views->assign(nl, std::vector<void*>(W, nullptr));
    if (nl == W) {
        for (int i = 0; i < nl; i++)
            for (int sidx = 0; sidx < W; sidx++) (*views)[i][sidx] = own[sidx];     // local rank index == global rank (init_all)
        return 0;
    }
    return for_each_rank(c, [&](int i) -> int {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        const int me = c->rank[i];
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
        std::vector<cudaIpcMemHandle_t> h(W);
        memset(h.data(), 0, sizeof(cudaIpcMemHandle_t) * W);
        if (!own[i] || cudaIpcGetMemHandle(&h[me], own[i]) != cudaSuccess) { (void)cudaGetLastError(); (*ok_local)[i] = 0; }
        unsigned char* d = nullptr;
        TRY(dev_alloc(ctx, (size_t)W * 64, (void**)&d));
        int r = 0;
        if (cudaMemcpyAsync(d + (size_t)me * 64, &h[me], 64, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) r = B200ZKP_ERR_CUDA;
        if (!r) {
            ncclResult_t nr = nccl_api().AllGather(d + (size_t)me * 64, d, 64, ncclChar, c->nc[i], ctx->stream);
            if (nr != ncclSuccess) r = nccl_fail(&ctx->err, "ncclAllGather", nr);
        }
        if (!r && (cudaMemcpyAsync(h.data(), d, (size_t)W * 64, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                   cudaStreamSynchronize(ctx->stream) != cudaSuccess)) r = B200ZKP_ERR_CUDA;
        dev_release(ctx, d, (size_t)W * 64);
        if (r) { if (r == B200ZKP_ERR_CUDA) { ctx->err = "CUDA error while exchanging IPC handles"; (void)cudaGetLastError(); } return r; }
        for (int sidx = 0; sidx < W; sidx++) {
            if (sidx == me) { (*views)[i][sidx] = own[i]; continue; }
            void* ptr = nullptr;
            if (!(*ok_local)[i] || cudaIpcOpenMemHandle(&ptr, h[sidx], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                (void)cudaGetLastError();
                (*ok_local)[i] = 0;
                continue;
            }
            (*views)[i][sidx] = ptr;
            (*imported)[i].push_back(ptr);
        }
        return 0;
    });
}

static void peer_close_imports(std::vector<void*>* v) {
    for (void* q : *v) if (q) cudaIpcCloseMemHandle(q);
    (void)cudaGetLastError();
    v->clear();
}

// flags of every rank, mapped into every rank; decides collectively whether the peer-memory exchange can be used at all
static int comm_peer_init(b200zkp_comm* c) {
    const int nl = c->n_local(), W = c->world;
    c->pr.assign(nl, PeerRank());
    c->peer_ok = false;
    const char* env = getenv("B200ZKP_PEER_EXCHANGE");
    if (env && env[0] == '0') c->peer_on = false;
    if (const char* pg = getenv("B200ZKP_PEER_GROUPS")) { const u32 v = (u32)strtoul(pg, nullptr, 10); if (v >= 1 && v <= 16) c->pull_groups = v; }
    if (const char* cb = getenv("B200ZKP_PEER_CHUNK_BYTES")) { const u64 v = strtoull(cb, nullptr, 10); if (v >= 8) c->chunk_bytes = v; }
    if (W < 2) return 0;
    std::vector<int> ok(nl, 1);
    std::vector<void*> own(nl, nullptr);
    for (int i = 0; i < nl; i++) {
        if (W > (int)peer::MAX_RANKS || (nl != W && nl != 1)) { ok[i] = 0; continue; }
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        if (nl == W) {
            // one process: direct peer access between every pair of devices
            for (int j = 0; j < nl && ok[i]; j++) {
                if (j == i) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, ctx->device, c->ctx[j]->device) != cudaSuccess || !can) { ok[i] = 0; break; }
                cudaError_t e = cudaDeviceEnablePeerAccess(c->ctx[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok[i] = 0;
                (void)cudaGetLastError();
            }
        }
        if (ok[i]) {
            // (a whole 2 MB block of its own: what CUDA IPC exports is the block, not the words)
            if (cudaMalloc((void**)&c->pr[i].flags, (size_t)2 << 20) != cudaSuccess ||
                cudaMemsetAsync(c->pr[i].flags, 0, (size_t)2 << 20, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess) { (void)cudaGetLastError(); ok[i] = 0; }
            own[i] = c->pr[i].flags;
        }
    }
    bool all = false;
    TRY(comm_all_ok(c, ok, &all));
    if (all) {
        std::vector<std::vector<void*>> views, imported(nl);
        TRY(comm_exchange_ptrs(c, own, &views, &imported, &ok));
        for (int i = 0; i < nl; i++) {
            for (int sidx = 0; sidx < W; sidx++) c->pr[i].peer_flags.p[sidx] = (u32*)views[i][sidx];
            c->pr[i].imported_flags = imported[i];
        }
        TRY(comm_all_ok(c, ok, &all));
    }
    if (all && c->one_process()) {
        for (int i = 0; i < nl && all; i++) {
            cudaSetDevice(c->ctx[i]->device);
            for (u32 e = 0; e <= peer::MAX_CHUNKS && all; e++) {
                cudaEvent_t ev = nullptr;
                if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); all = false; break; }
                if (e < peer::MAX_CHUNKS) c->ready_ev.push_back(ev); else c->done_ev.push_back(ev);
            }
        }
    }
    c->peer_ok = all;
    if (!all) {
        for (int i = 0; i < nl; i++) {
            cudaSetDevice(c->ctx[i]->device);
            peer_close_imports(&c->pr[i].imported_flags);
        }
        // the owners free only after every importer has closed
        std::vector<int> one(nl, 1);
        bool dummy;
        TRY(comm_all_ok(c, one, &dummy));
        for (int i = 0; i < nl; i++) {
            cudaSetDevice(c->ctx[i]->device);
            if (c->pr[i].flags) cudaFree(c->pr[i].flags);
            c->pr[i].flags = nullptr;
        }
        (void)cudaGetLastError();
    }
    return 0;
}

// the window of every rank holds at least `bytes` (collective; the same request on every rank).  Growing it is a full
// rendezvous: readers finish, importers close, owners reallocate, the new handles travel.
static int comm_ensure_window(b200zkp_comm* c, size_t bytes) {
    if (!c->peer_ok || c->pr.empty() || bytes <= c->pr[0].win_b) return 0;
    const int nl = c->n_local(), W = c->world;
    bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    std::vector<int> ok(nl, 1);
    // 1. nobody reads a window any more: every rank has seen `done` of the last commit from every peer
    TRY(for_each_rank(c, [&](int i) -> int {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        // (one process: every reader's stream is drained by this very loop before anything is freed)
        if (!c->one_process()) {
            peer::wait_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].flags, peer::DONE0, 1, (u32)W, (u32)c->rank[i], c->epoch, PEER_WAIT_TIMEOUT_NS);
            LAUNCH_CHECK(ctx);
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        peer_close_imports(&c->pr[i].imported_win);
        return 0;
    }));
    bool all = false;
    TRY(comm_all_ok(c, ok, &all));                 // barrier: every importer has closed
    std::vector<void*> own(nl, nullptr);
    for (int i = 0; i < nl; i++) {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        if (c->pr[i].win) cudaFree(c->pr[i].win);
        c->pr[i].win = nullptr; c->pr[i].win_b = 0;
        if (cudaMalloc((void**)&c->pr[i].win, bytes) != cudaSuccess) {
            (void)cudaGetLastError();
            pool_drop(ctx);
            if (cudaMalloc((void**)&c->pr[i].win, bytes) != cudaSuccess) { (void)cudaGetLastError(); c->pr[i].win = nullptr; ok[i] = 0; }
        }
        own[i] = c->pr[i].win;
    }
    std::vector<std::vector<void*>> views, imported(nl);
    TRY(comm_exchange_ptrs(c, own, &views, &imported, &ok));
    for (int i = 0; i < nl; i++) {
        c->pr[i].peer_win.assign(W, nullptr);
        for (int sidx = 0; sidx < W; sidx++) c->pr[i].peer_win[sidx] = (const u64*)views[i][sidx];
        c->pr[i].imported_win = imported[i];
        c->pr[i].win_b = bytes;
    }
    TRY(comm_all_ok(c, ok, &all));
    if (!all) {
        // (out of memory on some rank, or a mapping failed) fall back to the NCCL exchange for good
        for (int i = 0; i < nl; i++) { cudaSetDevice(c->ctx[i]->device); peer_close_imports(&c->pr[i].imported_win); }
        std::vector<int> one(nl, 1);
        bool dummy;
        TRY(comm_all_ok(c, one, &dummy));
        for (int i = 0; i < nl; i++) {
            cudaSetDevice(c->ctx[i]->device);
            if (c->pr[i].win) cudaFree(c->pr[i].win);
            c->pr[i].win = nullptr; c->pr[i].win_b = 0;
        }
        (void)cudaGetLastError();
        c->peer_ok = false;
    }
    return 0;
}

// before the communicator goes away: readers finish, importers close (barrier), owners free
static void comm_peer_shutdown(b200zkp_comm* c) {
    if (c->pr.empty()) return;
    const int nl = c->n_local();
    if (c->peer_ok) {
        (void)for_each_rank(c, [&](int i) -> int {
            b200zkp_ctx* ctx = c->ctx[i];
            Guard g(ctx);
            if (!c->one_process())
                peer::wait_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].flags, peer::DONE0, 1, (u32)c->world, (u32)c->rank[i], c->epoch, PEER_WAIT_TIMEOUT_NS);
            cudaStreamSynchronize(ctx->stream);
            peer_close_imports(&c->pr[i].imported_win);
            peer_close_imports(&c->pr[i].imported_flags);
            return 0;
        });
        std::vector<int> one(nl, 1);
        bool dummy;
        (void)comm_all_ok(c, one, &dummy);
    }
    for (int i = 0; i < nl; i++) {
        cudaSetDevice(c->ctx[i]->device);
        if (c->pr[i].win) cudaFree(c->pr[i].win);
        if (c->pr[i].flags) cudaFree(c->pr[i].flags);
    }
    for (size_t e = 0; e < c->ready_ev.size(); e++) { cudaSetDevice(c->ctx[e / peer::MAX_CHUNKS]->device); cudaEventDestroy(c->ready_ev[e]); }
    for (size_t e = 0; e < c->done_ev.size(); e++) { cudaSetDevice(c->ctx[e]->device); cudaEventDestroy(c->done_ev[e]); }
    c->ready_ev.clear(); c->done_ev.clear();
    (void)cudaGetLastError();
    c->pr.clear();
    c->peer_ok = false;
}

static int comm_finish_init(b200zkp_comm* c) {
    // side stream + events of every local rank
    for (int i = 0; i < c->n_local(); i++) {
        b200zkp_ctx* ctx = c->ctx[i];
        if (cudaSetDevice(ctx->device) != cudaSuccess) { (void)cudaGetLastError(); c->err = "cudaSetDevice failed"; return B200ZKP_ERR_CUDA; }
        cudaStream_t s = nullptr;
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi) != cudaSuccess) {
            (void)cudaGetLastError(); c->err = "cannot create the exchange stream"; return B200ZKP_ERR_CUDA;
        }
        c->xstream.push_back(s);
        cudaStream_t hs = nullptr;
        if (cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); c->err = "cannot create the hash stream"; return B200ZKP_ERR_CUDA; }
        c->hstream.push_back(hs);
        for (int e = 0; e < 2 + c->world; e++) {
            cudaEvent_t ev = nullptr;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); c->err = "cannot create an event"; return B200ZKP_ERR_CUDA; }
            c->ev.push_back(ev);
        }
    }
    return 0;
}

static bool valid_world(int w) { return w >= 1 && w <= 256 && (w & (w - 1)) == 0; }

extern "C" int b200zkp_comm_init_rank(b200zkp_ctx* ctx, const uint8_t id[B200ZKP_COMM_ID_BYTES], int rank, int world,
                                      b200zkp_comm** out) {
    if (!out) return B200ZKP_ERR_BAD_ARG;
    *out = nullptr;
    if (!ctx) return B200ZKP_ERR_BAD_ARG;
    b200zkp_comm* c = nullptr;
    int rc = 0;
    {
        Guard g(ctx);           // (released before the peer handshake, whose steps lock the ctx themselves)
        if (!id || !valid_world(world) || rank < 0 || rank >= world) BAD(ctx, "bad communicator shape (world must be a power of two)");
        if ((rc = nccl_ready(&ctx->err))) return rc;
        c = new (std::nothrow) b200zkp_comm();
        if (!c) return B200ZKP_ERR_OOM;
        c->world = world;
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        ncclComm_t nc = nullptr;
        ncclResult_t r = nccl_api().CommInitRank(&nc, world, u, rank);
        if (r != ncclSuccess) { rc = nccl_fail(&ctx->err, "ncclCommInitRank", r); delete c; return rc; }
        c->rank.push_back(rank); c->ctx.push_back(ctx); c->nc.push_back(nc);
        rc = comm_finish_init(c);
        if (rc) ctx->err = c->err;
    }
    if (!rc) rc = comm_peer_init(c);
    if (rc) { b200zkp_comm_destroy(c); return rc; }
    *out = c;
    return 0;
}

extern "C" int b200zkp_comm_init_all(b200zkp_ctx* const* ctxs, int n, b200zkp_comm** out) {
    if (!out) return B200ZKP_ERR_BAD_ARG;
    *out = nullptr;
    if (!ctxs || !valid_world(n)) return B200ZKP_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return B200ZKP_ERR_BAD_ARG;
        for (int j = 0; j < i; j++)
            if (ctxs[j] == ctxs[i] || ctxs[j]->device == ctxs[i]->device) { ctxs[0]->err = "init_all needs one ctx per distinct device"; return B200ZKP_ERR_BAD_ARG; }
    }
    if (int rc = nccl_ready(&ctxs[0]->err)) return rc;
    b200zkp_comm* c = new (std::nothrow) b200zkp_comm();
    if (!c) return B200ZKP_ERR_OOM;
    c->world = n;
    std::vector<int> devs(n);
    for (int i = 0; i < n; i++) { devs[i] = ctxs[i]->device; c->rank.push_back(i); c->ctx.push_back(ctxs[i]); }
    c->nc.assign(n, nullptr);
    ncclResult_t r = nccl_api().CommInitAll(c->nc.data(), n, devs.data());
    if (r != ncclSuccess) { int rc = nccl_fail(&ctxs[0]->err, "ncclCommInitAll", r); c->nc.clear(); delete c; return rc; }
    if (int rc = comm_finish_init(c)) { ctxs[0]->err = c->err; b200zkp_comm_destroy(c); return rc; }
    if (int rc = comm_peer_init(c)) { ctxs[0]->err = c->err; b200zkp_comm_destroy(c); return rc; }
    *out = c;
    return 0;
}

extern "C" void b200zkp_comm_destroy(b200zkp_comm* c) {
    if (!c) return;
    comm_peer_shutdown(c);
    for (int i = 0; i < c->n_local(); i++) {
        cudaSetDevice(c->ctx[i]->device);
        if (i < (int)c->xstream.size()) { cudaStreamSynchronize(c->xstream[i]); }
        cudaStreamSynchronize(c->ctx[i]->stream);
    }
    for (size_t i = 0; i < c->nc.size(); i++) if (c->nc[i]) nccl_api().CommDestroy(c->nc[i]);
    for (int i = 0; i < (int)c->xstream.size(); i++) { cudaSetDevice(c->ctx[i]->device); cudaStreamDestroy(c->xstream[i]); }
    for (int i = 0; i < (int)c->hstream.size(); i++) { cudaSetDevice(c->ctx[i]->device); cudaStreamSynchronize(c->hstream[i]); cudaStreamDestroy(c->hstream[i]); }
    for (size_t e = 0; e < c->ev.size(); e++) {
        cudaSetDevice(c->ctx[e / (2 + c->world)]->device);
        cudaEventDestroy(c->ev[e]);
    }
    (void)cudaGetLastError();
    delete c;
}

extern "C" const char* b200zkp_comm_last_error(const b200zkp_comm* c) { return c ? c->err.c_str() : "null comm"; }

extern "C" int b200zkp_comm_shape(const b200zkp_comm* c, int32_t shape[3]) {
    if (!c || !shape) return B200ZKP_ERR_BAD_ARG;
    shape[0] = c->world; shape[1] = c->n_local(); shape[2] = c->rank[0];
    return 0;
}

extern "C" int b200zkp_comm_set_exchange_group(b200zkp_comm* c, uint32_t peers_per_group) {
    if (!c) return B200ZKP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    c->peers_per_group = peers_per_group;
    return 0;
}

extern "C" int b200zkp_comm_set_peer_exchange(b200zkp_comm* c, int enabled) {
    if (!c) return B200ZKP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    c->peer_on = enabled != 0;
    return 0;
}

extern "C" int b200zkp_comm_peer_exchange(const b200zkp_comm* c) { return (c && c->peer_exchange()) ? 1 : 0; }

static void sharded_release(b200zkp_sharded* sh) {
    for (size_t i = 0; i < sh->r.size(); i++) {
        b200zkp_ctx* ctx = sh->comm->ctx[i];
        Guard g(ctx);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(sh->comm->xstream[i]);
        cudaStreamSynchronize(sh->comm->hstream[i]);
        ShardRank& s = sh->r[i];
        dev_release(ctx, s.coeffs_all, s.coeffs_b); dev_release(ctx, s.lde, s.lde_b); dev_release(ctx, s.digests, s.digests_b);
        dev_release(ctx, s.cap_local, s.cap_local_b); dev_release(ctx, s.cap, s.cap_b); dev_release(ctx, s.stage, s.stage_b);
        dev_release(ctx, s.gather, s.gather_b); dev_release(ctx, s.sponge, s.sponge_b);
    }
    delete sh;
}

extern "C" int b200zkp_sharded_create(b200zkp_comm* c, uint32_t n_log, uint32_t k, uint32_t rate_bits, uint32_t cap_height,
                                      b200zkp_sharded** out) {
    if (!out) return B200ZKP_ERR_BAD_ARG;
    *out = nullptr;
    if (!c) return B200ZKP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    const u32 G = (u32)c->world;
    if (k == 0) COMM_BAD(c, "empty polynomial batch");
    if (rate_bits > 8 || n_log + rate_bits > 32) COMM_BAD(c, "n_log + rate_bits exceeds two-adicity");
    if (cap_height > n_log + rate_bits) COMM_BAD(c, "cap_height exceeds log2(LDE size)");
    if (G > (1u << rate_bits) || G > (1u << cap_height)) COMM_BAD(c, "world size must be <= 2^rate_bits and <= 2^cap_height");
    b200zkp_sharded* sh = new (std::nothrow) b200zkp_sharded();
    if (!sh) return B200ZKP_ERR_OOM;
    sh->comm = c; sh->n_log = n_log; sh->k = k; sh->rate_bits = rate_bits; sh->cap_height = cap_height;
    sh->L = std::max(8u, G);
    sh->w = sh->L / G;
    sh->n_groups = (k + sh->L - 1) / sh->L;
    sh->kp = sh->n_groups * sh->w;
    sh->bpr = (1u << rate_bits) / G;
    sh->cpr = (1u << cap_height) / G;
    u32 g_log = 0;
    while ((1u << g_log) < G) g_log++;
    sh->cap_height_local = cap_height - g_log;
    const u64 n = (u64)1 << n_log;
    sh->N_local = (n << rate_bits) / G;
    sh->r.resize(c->n_local());
    int rc = for_each_rank(c, [&](int i) -> int {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        ShardRank& s = sh->r[i];
        s.coeffs_b = (size_t)sh->kp * G * n * 8;             // >= k columns; the tail stays zero
        s.lde_b = (size_t)k * sh->N_local * 8;
        s.digests_b = (size_t)2 * (sh->N_local - ((u64)1 << sh->cap_height_local)) * 32;
        s.cap_local_b = (size_t)32 << sh->cap_height_local;
        s.cap_b = (size_t)32 << cap_height;
        TRY(dev_alloc(ctx, s.coeffs_b, (void**)&s.coeffs_all));
        TRY(dev_alloc(ctx, s.lde_b, (void**)&s.lde));
        TRY(dev_alloc(ctx, s.digests_b, (void**)&s.digests));
        TRY(dev_alloc(ctx, s.cap_local_b, (void**)&s.cap_local));
        TRY(dev_alloc(ctx, s.cap_b, (void**)&s.cap));
        CUDA_TRY(ctx, cudaMemsetAsync(s.coeffs_all, 0, s.coeffs_b, ctx->stream));
        return 0;
    });
    if (!rc && c->peer_exchange()) rc = comm_ensure_window(c, (size_t)sh->kp * n * 8);
    if (rc) { sharded_release(sh); return rc; }
    *out = sh;
    return 0;
}

extern "C" void b200zkp_sharded_free(b200zkp_sharded* sh) {
    if (!sh) return;
    std::lock_guard<std::mutex> lk(sh->comm->mu);
    sharded_release(sh);
}

extern "C" int b200zkp_sharded_layout(const b200zkp_sharded* sh, int local, uint64_t lay[8]) {
    if (!sh || !lay || local < 0 || local >= sh->comm->n_local()) return B200ZKP_ERR_BAD_ARG;
    const u64 g = (u64)sh->comm->rank[local];
    lay[0] = sh->local_cols((u32)g);
    lay[1] = sh->L;
    lay[2] = sh->w;
    lay[3] = g * sh->bpr; lay[4] = (g + 1) * sh->bpr;
    lay[5] = sh->N_local;
    lay[6] = g * sh->cpr; lay[7] = (g + 1) * sh->cpr;
    return 0;
}

extern "C" int b200zkp_sharded_columns(const b200zkp_sharded* sh, int local, uint32_t* cols, uint32_t capacity, uint32_t* n_cols) {
    if (!sh || !n_cols || local < 0 || local >= sh->comm->n_local()) return B200ZKP_ERR_BAD_ARG;
    const u32 g = (u32)sh->comm->rank[local], cnt = sh->local_cols(g);
    *n_cols = cnt;
    if (cols) {
        if (capacity < cnt) return B200ZKP_ERR_BAD_ARG;
        for (u32 t = 0; t < cnt; t++) cols[t] = sh->global_col(g, t);
    }
    return 0;
}

// ---- ordering between the ranks of a peer exchange: flags in peer memory between processes (peer_kernels.cuh), CUDA events
//      plus a rendezvous of the rank threads inside one process
static int peer_wait_window_free(b200zkp_comm* c, int i, u32 epoch) {
    b200zkp_ctx* ctx = c->ctx[i];
    if (c->one_process()) {
        if (c->done_valid)
            for (int sidx = 0; sidx < c->n_local(); sidx++)
                if (sidx != i) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, c->done_ev[sidx], 0));
        return 0;
    }
    peer::wait_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].flags, peer::DONE0, 1, (u32)c->world, (u32)c->rank[i], epoch - 1, PEER_WAIT_TIMEOUT_NS);
    LAUNCH_CHECK(ctx);
    return 0;
}
// chunk j of this rank's window is written (stream order); returns when the stream also waits for chunk j of every peer
static int peer_publish_and_wait_chunk(b200zkp_comm* c, int i, u32 j, u32 epoch) {
    b200zkp_ctx* ctx = c->ctx[i];
    if (c->one_process()) {
        CUDA_TRY(ctx, cudaEventRecord(c->ready_ev[(size_t)i * peer::MAX_CHUNKS + j], ctx->stream));
        if (!c->hb.wait()) { ctx->err = "another rank of the communicator failed"; return B200ZKP_ERR_CUDA; }
        for (int sidx = 0; sidx < c->n_local(); sidx++)
            if (sidx != i) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, c->ready_ev[(size_t)sidx * peer::MAX_CHUNKS + j], 0));
        return 0;
    }
    const u32 G = (u32)c->world, g = (u32)c->rank[i];
    peer::signal_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].peer_flags, G, g, peer::READY0 + g * peer::MAX_CHUNKS + j, epoch);
    LAUNCH_CHECK(ctx);
    peer::wait_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].flags, peer::READY0 + j, peer::MAX_CHUNKS, G, g, epoch, PEER_WAIT_TIMEOUT_NS);
    LAUNCH_CHECK(ctx);
    return 0;
}
// this rank has read every window (stream order)
static int peer_publish_done(b200zkp_comm* c, int i, u32 epoch) {
    b200zkp_ctx* ctx = c->ctx[i];
    if (c->one_process()) {
        CUDA_TRY(ctx, cudaEventRecord(c->done_ev[i], ctx->stream));
        return 0;
    }
    peer::signal_kernel<<<1, 32, 0, ctx->stream>>>(c->pr[i].peer_flags, (u32)c->world, (u32)c->rank[i], peer::DONE0 + (u32)c->rank[i], epoch);
    LAUNCH_CHECK(ctx);
    return 0;
}

// where this rank's input columns live: packed (local column t at base + t * n) or, for the one-call form on the whole
// matrix, at their place in the k x n matrix (local column t = column global_col(g, t))
struct ShardInput { const u64* base; bool full_matrix; };

static int ensure_buffer(b200zkp_ctx* ctx, u64** buf, size_t* have, size_t want) {
    if (*have >= want) return 0;
    dev_release(ctx, *buf, *have);
    *buf = nullptr; *have = 0;
    TRY(dev_alloc(ctx, want, (void**)buf));
    *have = want;
    return 0;
}

// Everything one rank does for one commit; asynchronous (returns after enqueueing).
//
// The rank's columns are cut into the same chunks (whole column groups) on every rank; per chunk:
//   [upload ->] inverse transform into the rank's shard buffer -> (peer form) "ready" flag to every peer, wait for every
//   peer's flag -> gather + coset transforms of the chunk's columns of ALL ranks [-> absorb them into the leaf sponges].
// Peer form (the default inside one NVLink domain): the shard buffer is the rank's exchange window and the gather happens
// inside the first pass of the coset transforms, which reads the coefficient tiles straight from the owners' windows over NVLink
// and files them in the local coefficient matrix on the way (ntc::ct_pull_kernel).  With host inputs the upload of chunk j + 1
// runs under the transforms and the hashing of chunk j (resumable sponge), so only the first chunk's upload is exposed.
// NCCL form (fallback): the shards travel by ncclSend / ncclRecv into a local gather buffer after the last inverse transform,
// and the same kernels then read them from there.
static int sharded_commit_rank(b200zkp_sharded* sh, int i, ShardInput in, int on_device, int is_coeffs, bool peer, u32 epoch) {
    b200zkp_comm* c = sh->comm;
    b200zkp_ctx* ctx = c->ctx[i];
    NcclApi& nc = nccl_api();
    Guard guard(ctx);
    ShardRank& s = sh->r[i];
    const u32 G = (u32)c->world, g = (u32)c->rank[i];
    const u64 n = (u64)1 << sh->n_log;
    const u32 L = sh->L, w = sh->w, kp = sh->kp, k = sh->k, kl = sh->local_cols(g);
    cudaStream_t main_s = ctx->stream, side_s = c->xstream[i], hash_s = c->hstream[i];
    if (kl && !in.base) BAD(ctx, "null input shard");
    if (on_device && in.full_matrix) BAD(ctx, "internal: device inputs are packed shards");

    // ---- the rank's shard buffer and where the shards of the others are read from
    u64* own_out = nullptr;
    const u64* src[ntc::MAX_SRC] = {};
    const bool few = G <= (u32)ntc::MAX_SRC;
    if (peer) {
        PeerRank& pr = c->pr[i];
        if (!few || !pr.win || pr.win_b < (size_t)kp * n * 8 || pr.peer_win.size() != G) BAD(ctx, "internal: exchange window missing");
        own_out = pr.win;
        for (u32 q = 0; q < G; q++) src[q] = pr.peer_win[q];
    } else {
        TRY(ensure_buffer(ctx, &s.gather, &s.gather_b, (size_t)G * kp * n * 8));
        own_out = s.gather + (u64)g * kp * n;
        if (few) for (u32 q = 0; q < G; q++) src[q] = s.gather + (u64)q * kp * n;
    }
    const bool covered = ctx->ntt_ct && ntc::covers(sh->n_log);
    // the gather inside the first pass wants tile rows of 128 bytes and more over NVLink (first passes of <= 7 bits) unless the
    // upload of host inputs hides the transfer anyway; from the local gather buffer it always pays
    const bool fused = covered && few && sh->bpr <= (u32)ntc::MAX_LOOP_BLOCKS &&
                       (!peer || !on_device || ntc::first_pass_bits(sh->n_log) <= 7 || getenv("B200ZKP_PEER_FUSED_ALWAYS"));
    // tables first: their builders synchronise the stream, and nothing of this commit may be waiting on a peer by then
    if (covered) {
        b200zkp_ctx::ZTables z;
        TRY(get_ztab_lde(ctx, sh->n_log, sh->rate_bits, &z));
        if (!is_coeffs) TRY(get_ztab_inv(ctx, sh->n_log, &z));
    }

    // ---- chunks of whole column groups: the same cut on every rank (host inputs: >= chunk_bytes per upload)
    u32 m = sh->n_groups;
    if (!on_device) {
        m = (u32)std::max<u64>(1, (c->chunk_bytes + (u64)w * n * 8 - 1) / ((u64)w * n * 8));
        while ((sh->n_groups + m - 1) / m > peer::MAX_CHUNKS) m++;
    }
    const u32 C = (sh->n_groups + m - 1) / m;
    merkle::TreeShape shape;
    TRY(merkle_shape(ctx, sh->N_local, sh->cap_height_local, s.lde, s.digests, s.cap_local, &shape));
    // host inputs in the peer form: the leaves are hashed chunk by chunk as their columns arrive
    const bool resumable = peer && !on_device && C > 1 && k > 4 && sh->N_local > COOP_MAX_NODES && !getenv("B200ZKP_NO_RESUMABLE_SPONGE");
    if (resumable) TRY(ensure_buffer(ctx, &s.sponge, &s.sponge_b, (size_t)poseidon::WIDTH * sh->N_local * 8));

    std::vector<cudaEvent_t> up(C, nullptr);
    if (!on_device && kl) {
        TRY(ensure_buffer(ctx, &s.stage, &s.stage_b, (size_t)kp * n * 8));
        if (!ctx->stream2) {
            int lo = 0, hi = 0;
            CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, hi));
        }
        cudaEvent_t e_free;
        TRY(get_sync_event(ctx, 0, &e_free));
        CUDA_TRY(ctx, cudaEventRecord(e_free, main_s));                  // the staging buffer may still feed the previous step
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, e_free, 0));
        for (u32 j = 0; j < C; j++) {
            const u32 t0 = std::min(kl, j * m * w), t1 = std::min(kl, (j + 1) * m * w);
            TRY(get_sync_event(ctx, 1 + j, &up[j]));
            if (t1 > t0 && !in.full_matrix)
                CUDA_TRY(ctx, cudaMemcpyAsync(s.stage + (u64)t0 * n, in.base + (u64)t0 * n, (size_t)(t1 - t0) * n * 8, cudaMemcpyHostToDevice, ctx->stream2));
            if (t1 > t0 && in.full_matrix)
                for (u32 t = t0; t < t1; t += w) {       // the w local columns of a group are neighbours in the matrix
                    const u32 cnt = std::min(w, t1 - t);
                    CUDA_TRY(ctx, cudaMemcpyAsync(s.stage + (u64)t * n, in.base + (u64)sh->global_col(g, t) * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice, ctx->stream2));
                }
            CUDA_TRY(ctx, cudaEventRecord(up[j], ctx->stream2));
        }
    }
    const u64* in_cols = on_device ? in.base : s.stage;
    u32 n_ev = 0;
    bool side_used = false;
    std::unique_ptr<StageTimer> lde_timer;                // device inputs: gather + coset transforms as one span of the main stream
    auto next_event = [&](cudaEvent_t* e) -> int { return get_sync_event(ctx, 24 + (n_ev++ & 63), e); };
    auto join_side = [&]() -> int {
        if (!side_used) return 0;
        cudaEvent_t e_join;
        TRY(next_event(&e_join));
        CUDA_TRY(ctx, cudaEventRecord(e_join, side_s));
        CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, e_join, 0));
        side_used = false;
        return 0;
    };

    // gather + coset transforms of the columns [pa, pb) (pa a multiple of L) of every rank, read from src[]
    auto gather_lde = [&](u32 pa, u32 pb, bool split) -> int {
        const u32 b0 = g * sh->bpr, b1 = (g + 1) * sh->bpr;
        const u32 grp0 = pa / L, grp1 = (pb + L - 1) / L;
        if (fused) {
            // device inputs arrive as one chunk: cut it into column ranges, so that the gather pass of range t + 1 (NVLink
            // bound, launched two CTAs per SM wide) runs beside the in-place passes of range t on the side stream
            const u32 n_rng = split ? std::min(c->pull_groups, std::max(1u, (pb - pa) / 16)) : 1u;
            const u32 rw = (pb - pa + n_rng - 1) / n_rng;
            for (u32 ra = pa; ra < pb; ra += rw) {
                const u32 rb = std::min(pb, ra + rw);
                ntc::ColumnSet cs;
                cs.run = w; cs.period = L; cs.col0 = ra; cs.limit = k; cs.count = rb - ra; cs.sel = 0; cs.pull = true;
                for (u32 q = 0; q < G; q++) cs.src[q] = src[q];
                cs.src_col_stride = n; cs.copy_out = s.coeffs_all; cs.copy_col_stride = n;
                cudaEvent_t e_first;
                TRY(next_event(&e_first));
                if (!lde_timer && on_device) lde_timer.reset(new StageTimer(ctx, B200ZKP_STAGE_LDE));
                int rc = dev_lde_cols_locked(ctx, cs, nullptr, n, s.lde, sh->N_local, sh->n_log, sh->rate_bits, b0, b1, side_s, e_first, /*pull_ctas_per_sm=*/2);
                if (rc == B200ZKP_ERR_UNSUPPORTED) BAD(ctx, "internal: gather transform refused a covered shape");
                TRY(rc);
                side_used = true;
            }
            return 0;
        }
        // copies of rank q's columns of the range into the coefficient matrix: one run of <= w columns per group
        auto copy_shard = [&](u32 q, cudaStream_t st) -> int {
            if (!few && !peer) return 0;      // (more ranks than sources: see below)
            for (u32 grp = grp0; grp < grp1; grp++) {
                const u32 first = grp * L + q * w;
                if (first >= k) break;
                const u32 cnt = std::min(w, k - first);
                CUDA_TRY(ctx, cudaMemcpyAsync(s.coeffs_all + (u64)first * n, src[q] + (u64)grp * w * n, (size_t)cnt * n * 8, cudaMemcpyDefault, st));
            }
            return 0;
        };
        if (covered && few && sh->bpr <= (u32)ntc::MAX_LOOP_BLOCKS) {
            // first passes of 8 bits (tile rows of 64 bytes are too short for NVLink) with nothing to hide the transfer: the copy
            // engine pulls shard after shard on the side stream, own shard first, and the coset transforms of a shard run on
            // the main stream while the next one is in flight
            cudaEvent_t e_ready;
            TRY(next_event(&e_ready));
            CUDA_TRY(ctx, cudaEventRecord(e_ready, main_s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(side_s, e_ready, 0));
            for (u32 d = 0; d < G; d++) {
                const u32 q = (g + d) % G;
                const u32 cnt = std::min(sh->local_cols(q), grp1 * w) - std::min(sh->local_cols(q), grp0 * w);
                if (!cnt) continue;
                cudaEvent_t e_here;
                TRY(next_event(&e_here));
                TRY(copy_shard(q, side_s));
                CUDA_TRY(ctx, cudaEventRecord(e_here, side_s));
                CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, e_here, 0));
                ntc::ColumnSet cs;
                cs.run = w; cs.period = L; cs.col0 = pa; cs.limit = k; cs.count = (grp1 - grp0) * w; cs.sel = 1; cs.src_rank = q; cs.pull = false;
                int rc = dev_lde_cols_locked(ctx, cs, s.coeffs_all, n, s.lde, sh->N_local, sh->n_log, sh->rate_bits, b0, b1);
                if (rc == B200ZKP_ERR_UNSUPPORTED) BAD(ctx, "internal: shard transform refused a covered shape");
                TRY(rc);
            }
            return 0;
        }
        // one-pass transforms and wide blow-ups: plain copies, then the transforms of the whole range
        for (u32 q = 0; q < G; q++) {
            if (few || peer) { TRY(copy_shard(q, main_s)); continue; }
            for (u32 grp = grp0; grp < grp1; grp++) {          // (NCCL form with more ranks than gather sources)
                const u32 first = grp * L + q * w;
                if (first >= k) break;
                CUDA_TRY(ctx, cudaMemcpyAsync(s.coeffs_all + (u64)first * n, s.gather + ((u64)q * kp + (u64)grp * w) * n,
                                              (size_t)std::min(w, k - first) * n * 8, cudaMemcpyDeviceToDevice, main_s));
            }
        }
        return dev_lde_locked(ctx, s.coeffs_all + (u64)pa * n, n, s.lde + (u64)pa * sh->N_local, sh->N_local, sh->n_log, pb - pa,
                              sh->rate_bits, b0, b1);
    };

    // ---- the chunks
    if (peer) TRY(peer_wait_window_free(c, i, epoch));   // every peer has read the previous commit out of the window
    for (u32 j = 0; j < C; j++) {
        const u32 grp0 = j * m, grp1 = std::min(sh->n_groups, (j + 1) * m);
        const u32 t0 = std::min(kl, grp0 * w), t1 = std::min(kl, grp1 * w);
        const u32 pa = grp0 * L, pb = std::min(k, grp1 * L);
        if (t1 > t0) {
            if (!on_device) CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, up[j], 0));
            // scratch of the inverse transform: LDE columns of this very chunk, not written before its coset transforms
            if (is_coeffs) TRY(launch_canon_copy(ctx, in_cols + (u64)t0 * n, own_out + (u64)t0 * n, (u64)(t1 - t0) * n));
            else TRY(dev_intt_locked(ctx, in_cols + (u64)t0 * n, n, own_out + (u64)t0 * n, n, s.lde + (u64)pa * sh->N_local, sh->n_log, t1 - t0));
        }
        if (!peer) continue;
        TRY(peer_publish_and_wait_chunk(c, i, j, epoch));
        TRY(gather_lde(pa, pb, /*split=*/on_device != 0));
        if (resumable) {
            // the chunk's columns are absorbed on the hash stream as soon as their last pass is through; the main stream goes on
            // with the next chunk (its gather pass waits on NVLink more than it computes: the two kernels share the SMs)
            cudaEvent_t e_lde;
            TRY(next_event(&e_lde));
            CUDA_TRY(ctx, cudaEventRecord(e_lde, side_used ? side_s : main_s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(hash_s, e_lde, 0));
            const unsigned threads = hash_block_threads(sh->N_local);
            merkle::leaf_absorb_kernel<<<(unsigned)((sh->N_local + threads - 1) / threads), threads, 0, hash_s>>>(
                s.lde, sh->N_local, k, pa, pb, sh->N_local, shape, s.sponge, sh->N_local, s.digests, s.cap_local);
            LAUNCH_CHECK(ctx);
        }
    }
    if (resumable) {
        cudaEvent_t e_hash;
        TRY(next_event(&e_hash));
        CUDA_TRY(ctx, cudaEventRecord(e_hash, hash_s));
        CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, e_hash, 0));
    }
    if (peer) {
        TRY(peer_publish_done(c, i, epoch));             // every window has been read by this rank
    } else {
        // NCCL form: the shards travel in point-to-point groups on the exchange stream, then everything is local
        cudaEvent_t* ev = &c->ev[(size_t)i * (2 + G)];
        CUDA_TRY(ctx, cudaEventRecord(ev[0], main_s));
        CUDA_TRY(ctx, cudaStreamWaitEvent(side_s, ev[0], 0));
        const u32 per_group = c->peers_per_group ? c->peers_per_group : (G - 1);
        const size_t count = (size_t)kp * n;
        for (u32 d0 = 1; d0 < G; d0 += per_group) {
            const u32 d1 = std::min(G, d0 + per_group);
            NCCL_TRY(&ctx->err, nc.GroupStart());
            for (u32 d = d0; d < d1; d++) {
                const u32 to = (g + d) % G, from = (g + G - d) % G;
                NCCL_TRY(&ctx->err, nc.Send(own_out, count, ncclUint64, (int)to, c->nc[i], side_s));
                NCCL_TRY(&ctx->err, nc.Recv(s.gather + (u64)from * kp * n, count, ncclUint64, (int)from, c->nc[i], side_s));
            }
            NCCL_TRY(&ctx->err, nc.GroupEnd());
            ctx->launches++;
        }
        CUDA_TRY(ctx, cudaEventRecord(ev[1], side_s));
        CUDA_TRY(ctx, cudaStreamWaitEvent(main_s, ev[1], 0));
        TRY(gather_lde(0, k, /*split=*/true));
    }
    TRY(join_side());
    lde_timer.reset();

    // ---- the rank's leaves -> digests of its cap subtrees; the cap
    if (resumable) TRY(launch_tree_levels(ctx, sh->N_local, shape, s.digests, s.cap_local));
    else TRY(dev_merkle_locked(ctx, s.lde, /*row_stride=*/1, /*col_stride=*/sh->N_local, k, sh->N_local, sh->cap_height_local,
                               s.digests, s.cap_local));
    if (G > 1) {
        NCCL_TRY(&ctx->err, nc.AllGather(s.cap_local, s.cap, s.cap_local_b / 8, ncclUint64, c->nc[i], main_s));
        ctx->launches++;
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(s.cap, s.cap_local, s.cap_b, cudaMemcpyDeviceToDevice, main_s));
    }
    return 0;
}

// a wait of this rank's commits gave up (a peer never raised its flag): the results are not valid
static int peer_check_error(b200zkp_comm* c, int i) {
    if (!c->peer_ok || !c->pr[i].flags) return 0;
    b200zkp_ctx* ctx = c->ctx[i];
    u32 e = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&e, c->pr[i].flags + peer::ERROR_WORD, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (e) { ctx->err = "peer exchange timed out (a rank of the communicator did not take part in the commit)"; return B200ZKP_ERR_NCCL; }
    return 0;
}

static int sharded_commit_locked(b200zkp_sharded* sh, const u64* const* inputs, bool full_matrix, int on_device, int is_coeffs, u64* cap_out) {
    b200zkp_comm* c = sh->comm;
    if (!inputs) COMM_BAD(c, "null inputs");
    bool use_peer = c->peer_exchange() && c->world > 1 && c->world <= (int)ntc::MAX_SRC;
    if (use_peer) {
        TRY(comm_ensure_window(c, (size_t)sh->kp * ((size_t)8 << sh->n_log)));        // (a no-op unless the exchange was switched on after create)
        use_peer = c->peer_exchange();
    }
    const u32 epoch = use_peer ? ++c->epoch : 0;
    c->hb.reset(c->n_local());
    int rc = for_each_rank(c, [&](int i) -> int {
        const int r = sharded_commit_rank(sh, i, ShardInput{inputs[i], full_matrix}, on_device, is_coeffs, use_peer, epoch);
        if (r) c->hb.abort();          // the other rank threads must not wait for this one
        return r;
    });
    if (use_peer && c->one_process()) c->done_valid = rc == 0;
    TRY(rc);
    if (cap_out) {
        TRY(for_each_rank(c, [&](int i) -> int {
            b200zkp_ctx* ctx = c->ctx[i];
            Guard g(ctx);
            if (i == 0) TRY(d2h(ctx, cap_out, sh->r[0].cap, sh->r[0].cap_b));
            else CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            return use_peer ? peer_check_error(c, i) : 0;
        }));
    }
    return 0;
}

extern "C" int b200zkp_sharded_commit(b200zkp_sharded* sh, const uint64_t* const* inputs, int inputs_on_device, int is_coeffs,
                                      uint64_t* cap_out) {
    if (!sh) return B200ZKP_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(sh->comm->mu);
    return sharded_commit_locked(sh, (const u64* const*)inputs, /*full_matrix=*/false, inputs_on_device, is_coeffs, (u64*)cap_out);
}

extern "C" int b200zkp_sharded_commit_from_values(b200zkp_comm* c, const uint64_t* values, uint32_t n_log, uint32_t k,
                                                  uint32_t rate_bits, uint32_t cap_height, uint64_t* cap_out,
                                                  b200zkp_sharded** out) {
    if (!c || !out) return B200ZKP_ERR_BAD_ARG;
    *out = nullptr;
    if (!values) { std::lock_guard<std::mutex> lk(c->mu); COMM_BAD(c, "null values"); }
    b200zkp_sharded* sh = nullptr;
    TRY(b200zkp_sharded_create(c, n_log, k, rate_bits, cap_height, &sh));
    std::lock_guard<std::mutex> lk(c->mu);
    // every local rank picks its columns out of the one matrix
    std::vector<const u64*> in(c->n_local(), (const u64*)values);
    int rc = sharded_commit_locked(sh, in.data(), /*full_matrix=*/true, /*on_device=*/0, /*is_coeffs=*/0, (u64*)cap_out);
    if (!rc && !cap_out)   // the caller's host matrix must stay valid only for the duration of this call
        rc = for_each_rank(c, [&](int i) -> int { Guard g(c->ctx[i]); CUDA_TRY(c->ctx[i], cudaStreamSynchronize(c->ctx[i]->stream)); return 0; });
    if (rc) { sharded_release(sh); return rc; }
    *out = sh;
    return 0;
}

extern "C" int b200zkp_sharded_synchronize(b200zkp_sharded* sh) {
    if (!sh) return B200ZKP_ERR_BAD_ARG;
    b200zkp_comm* c = sh->comm;
    std::lock_guard<std::mutex> lk(c->mu);
    return for_each_rank(c, [&](int i) -> int {
        Guard g(c->ctx[i]);
        CUDA_TRY(c->ctx[i], cudaStreamSynchronize(c->ctx[i]->stream));
        return peer_check_error(c, i);
    });
}

extern "C" int b200zkp_sharded_device_ptrs(b200zkp_sharded* sh, int local, const uint64_t** coeffs, const uint64_t** lde,
                                           const uint64_t** digests, const uint64_t** cap) {
    if (!sh || local < 0 || local >= sh->comm->n_local()) return B200ZKP_ERR_BAD_ARG;
    const ShardRank& s = sh->r[local];
    if (coeffs) *coeffs = (const uint64_t*)s.coeffs_all;
    if (lde) *lde = (const uint64_t*)s.lde;
    if (digests) *digests = (const uint64_t*)s.digests;
    if (cap) *cap = (const uint64_t*)s.cap;
    return 0;
}

extern "C" int b200zkp_sharded_rows(b200zkp_sharded* sh, const uint64_t* idx, uint64_t n_idx, uint64_t* rows,
                                    uint64_t* siblings) {
    if (!sh) return B200ZKP_ERR_BAD_ARG;
    b200zkp_comm* c = sh->comm;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!n_idx) return 0;
    if (!idx) COMM_BAD(c, "null index list");
    const u64 N = sh->N_local * (u64)c->world;
    for (u64 q = 0; q < n_idx; q++) if (idx[q] >= N) COMM_BAD(c, "leaf index out of range");
    u32 lg = 0;
    while (((u64)1 << lg) < sh->N_local) lg++;
    const u32 depth = lg - sh->cap_height_local;
    const u32 k = sh->k;
    const bool multi_process = c->n_local() < c->world;
    // every local rank answers the indices it owns (compacted list -> gather -> scatter on the host); with one process
    // per GPU the answers of the other ranks arrive through one all-reduce (sum with zeros) of the small result buffers
    std::vector<std::vector<u64>> lrows(c->n_local()), lsib(c->n_local());
    std::vector<std::vector<u64>> lq(c->n_local());
    TRY(for_each_rank(c, [&](int i) -> int {
        b200zkp_ctx* ctx = c->ctx[i];
        Guard g(ctx);
        const u64 rk = (u64)c->rank[i];
        std::vector<u64> local_idx;
        for (u64 q = 0; q < n_idx; q++)
            if (idx[q] / sh->N_local == rk) { lq[i].push_back(q); local_idx.push_back(idx[q] % sh->N_local); }
        if (local_idx.empty()) return 0;
        if (rows) lrows[i].resize(local_idx.size() * k);
        if (siblings && depth) lsib[i].resize(local_idx.size() * depth * 4);
        return gather_locked(ctx, sh->r[i].lde, sh->N_local, k, sh->r[i].digests, depth, local_idx.data(), local_idx.size(),
                             rows ? lrows[i].data() : nullptr, (siblings && depth) ? lsib[i].data() : nullptr);
    }));
    if (rows) memset(rows, 0, (size_t)n_idx * k * 8);
    if (siblings && depth) memset(siblings, 0, (size_t)n_idx * depth * 32);
    for (int i = 0; i < c->n_local(); i++)
        for (size_t j = 0; j < lq[i].size(); j++) {
            if (rows) memcpy(rows + lq[i][j] * k, lrows[i].data() + j * k, (size_t)k * 8);
            if (siblings && depth) memcpy(siblings + lq[i][j] * depth * 4, lsib[i].data() + j * depth * 4, (size_t)depth * 32);
        }
    if (multi_process) {
        b200zkp_ctx* ctx = c->ctx[0];
        Guard g(ctx);
        NcclApi& nc = nccl_api();
        const size_t rows_w = rows ? (size_t)n_idx * k : 0, sib_w = (siblings && depth) ? (size_t)n_idx * depth * 4 : 0;
        const size_t total = rows_w + sib_w;
        if (total) {
            void* d = nullptr;
            TRY(dev_alloc(ctx, total * 8, &d));
            int rc = 0;
            if (rows_w) rc = h2d(ctx, d, rows, rows_w * 8);
            if (!rc && sib_w) rc = h2d(ctx, (u64*)d + rows_w, siblings, sib_w * 8);
            if (!rc) {
                ncclResult_t r = nc.AllReduce(d, d, total, ncclUint64, ncclSum, c->nc[0], ctx->stream);
                if (r != ncclSuccess) rc = nccl_fail(&ctx->err, "ncclAllReduce", r);
                ctx->launches++;
            }
            if (!rc && rows_w) rc = d2h(ctx, rows, d, rows_w * 8);
            if (!rc && sib_w) rc = d2h(ctx, siblings, (u64*)d + rows_w, sib_w * 8);
            if (rc) cudaStreamSynchronize(ctx->stream);
            dev_release(ctx, d, total * 8);
            if (rc) { c->err = ctx->err; return rc; }
        }
    }
    return 0;
}
