// Cross-GPU ordering for the partitioned commitment (sharded.inl): epoch-stamped flags in peer memory.
//
// Replaces nothing in plonky2 (one process, one memory); it is the synchronisation of north_star's multi-GPU form of
// PolynomialBatch::from_values, reached from the same prove() call sites
// (/root/reference/src/rollup/circuits/mod.rs:1247, src/transaction/circuits/mod.rs:453).
//
// Every rank owns one small flag array in device memory that every other rank of the NVLink domain can write (peer mapping
// or CUDA IPC).  A producer finishes its kernels, then a one-warp kernel stores the commit's epoch into its slot of every
// peer's array (release at system scope: the data the flag announces sits in the producer's memory and is read from there
// over NVLink, so it only has to be ordered before the flag).  A consumer's stream holds a one-warp kernel that spins on its
// OWN array (local memory, acquire at system scope) before the kernels that read the producer's exchange window.  The spin
// is bounded: after `timeout_ns` the waiter raises the error word and lets the stream go on, so a rank that died cannot hang
// the GPUs of the others (the commit then reports the error instead of a cap).
#pragma once
#include <cstdint>

namespace peer {

typedef unsigned int u32;
typedef unsigned long long u64;

static constexpr u32 MAX_RANKS = 8;       // ranks of one NVLink domain (= ntc::MAX_SRC)
static constexpr u32 MAX_CHUNKS = 16;     // column chunks of one commit (upload / transform / gather pipeline)
// flag words of one rank: ready[s][j] = epoch when chunk j of rank s's coefficients is in s's window; done[s] = epoch when
// rank s has finished reading MY window; error = 1 after a timed-out wait
static constexpr u32 READY0 = 0, DONE0 = MAX_RANKS * MAX_CHUNKS, ERROR_WORD = DONE0 + MAX_RANKS, FLAG_WORDS = ERROR_WORD + 1;

struct PeerFlags { u32* p[MAX_RANKS]; };

#ifndef B200ZKP_HOST_EMU
__device__ __forceinline__ void st_release_sys(u32* p, u32 v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ u32 ld_acquire_sys(const u32* p) {
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// lane t stores `value` into word `slot` of rank t's flags (every rank but `me`)
__global__ void signal_kernel(const PeerFlags peers, u32 world, u32 me, u32 slot, u32 value) {
    const u32 t = threadIdx.x;
    if (t >= world || t == me) return;
    __threadfence_system();
    st_release_sys(peers.p[t] + slot, value);
}

// lane t waits until word first_slot + t * stride of this rank's flags has reached `value` (every lane but `skip`)
__global__ void wait_kernel(u32* flags, u32 first_slot, u32 stride, u32 count, u32 skip, u32 value, u64 timeout_ns) {
    const u32 t = threadIdx.x;
    if (t >= count || t == skip) return;
    const u32* f = flags + first_slot + t * stride;
    const u64 t0 = global_ns();
    u32 spins = 0;
    while ((int)(ld_acquire_sys(f) - value) < 0) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) { atomicExch(flags + ERROR_WORD, 1u); break; }
    }
}
#endif

}  // namespace peer
