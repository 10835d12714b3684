#!/usr/bin/env python3
"""bench.py — polynomial-commitment throughput (BASELINE.json metric: LDE+Merkle commit cells/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one PolynomialBatch::from_values over BASELINE.json configs[2]: 2^20 rows x 135 columns, rate_bits 3,
Poseidon Merkle cap_height 4 (cell = one input element).  Inputs are synthetic (SplitMix64 mod p).
  value        whole-job cells/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e          same metric through the host-buffer C ABI call (b200zkp_commit_from_values): pinned host values
               -> H2D -> commit -> cap D2H, every step
  roofline     dominant kernel (Poseidon leaf hash): algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline the multithreaded CPU restatement of plonky2's path (oracle/cpu_baseline.c) on this box's cores,
               on a bounded sample of the same workload (reported baseline, not the target)
N > 1: ONE commitment partitioned over N GPUs (strong scaling): column-sharded iNTT, NCCL all-gather of the
coefficients, leaf-range-sharded LDE + hashing + cap subtrees, NCCL all-gather of the cap.
--impl reference: times the CPU restatement only (the reference's Rust prover cannot be built here: DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LOG, K, RATE_BITS, CAP_HEIGHT = 20, 135, 3, 4
METRIC = "LDE+Merkle commit cells/s"
UNIT = "cells/s"


def workload_name(n_log=N_LOG, k=K):
    return f"synthetic commitment 2^{n_log} rows x {k} columns, rate_bits={RATE_BITS}, Poseidon Merkle cap_height={CAP_HEIGHT}"


def algorithmic_bytes(n_log, k, r=RATE_BITS, h=CAP_HEIGHT):
    n, N = 1 << n_log, 1 << (n_log + r)
    return 8 * n * k + 8 * n * k + 8 * N * k + 64 * (N - (1 << h)) + 32 * (1 << h)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 8] or [r for (_, r) in self.rows if len(r) >= 8]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_commit_rate(n_log, k, steps, warmup):
    """cells/s of the multithreaded CPU restatement on a 2^n_log x k sample; returns (cells/s, ms/step, stage s)."""
    from oracle import oracle as O
    v = O.synthetic_values(k, 1 << n_log)
    for _ in range(warmup):
        O.baseline_commit(v, RATE_BITS, CAP_HEIGHT)
    ts, stages = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        _, st = O.baseline_commit(v, RATE_BITS, CAP_HEIGHT)
        ts.append(time.perf_counter() - t0)
        stages = st
    dt = sum(ts) / len(ts)
    return (k << n_log) / dt, dt * 1e3, stages


def cpu_baseline_sampled():
    """~10-30 s of CPU work: grow the sample until one commit takes >= 4 s (or 2^18 rows)."""
    from oracle import oracle as O
    cores = O.baseline_threads()
    n_log = 14
    rate, ms, stages = cpu_commit_rate(n_log, K, 1, 1)
    while ms < 4000 and n_log < 18:
        n_log += 2
        rate, ms, stages = cpu_commit_rate(n_log, K, 1, 0)
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"one from_values on 2^{n_log} x {K} (1/{1 << (N_LOG - n_log)} of the rows of the workload), "
                      f"{ms:.0f} ms; C restatement of plonky2's CPU path with OpenMP (oracle/cpu_baseline.c), "
                      f"stage s ifft/lde/transpose/merkle = " + "/".join(f"{s:.2f}" for s in stages[:4])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    n_log = int(os.environ.get("B200ZKP_REF_SAMPLE_LOG", "16"))
    rate, ms, stages = cpu_commit_rate(n_log, K, args.steps, max(args.warmup, 1))
    cores = O.baseline_threads()
    sample = (f"each step = one from_values on 2^{n_log} x {K} (1/{1 << (N_LOG - n_log)} of the rows), all {cores} host "
              f"threads; C restatement of plonky2 @ f99ed9c's CPU path (the Rust prover cannot be built here)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 (Goldilocks field)", "data": "synthetic",
        "config": {"workload": workload_name(), "sample_n_log": n_log, "k": K, "rate_bits": RATE_BITS, "cap_height": CAP_HEIGHT},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import intmax_zkp_core_b200 as z
    from intmax_zkp_core_b200 import device as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 through torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_log = int(os.environ.get("B200ZKP_BENCH_N_LOG", str(N_LOG)))
    k = int(os.environ.get("B200ZKP_BENCH_K", str(K)))
    n, N = 1 << n_log, 1 << (n_log + RATE_BITS)
    cells = n * k

    ctx = D.torch_context(local_rank)          # enqueues on torch's current stream: torch.cuda.Event sees our kernels
    if os.environ.get("B200ZKP_NO_OVERLAP"):
        ctx.set_overlap(False)
    lay = D.shard_layout(n_log, k, RATE_BITS, CAP_HEIGHT, rank, world)

    # synthetic input: v[c][i] = splitmix64(c*n + i) mod p; each rank generates only its column shard
    def synth(col_begin, col_end, pad_to):
        idx = np.arange(col_begin * n, col_end * n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            zed = (idx + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
            zed = (zed ^ (zed >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            zed = (zed ^ (zed >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            zed = zed ^ (zed >> np.uint64(31))
        zed = np.where(zed >= np.uint64(z.plonky2.P), zed - np.uint64(z.plonky2.P), zed).reshape(col_end - col_begin, n)
        host = torch.zeros((pad_to, n), dtype=torch.int64).pin_memory()
        host[:col_end - col_begin].copy_(torch.from_numpy(zed.view(np.int64)))
        return host

    host_vals = synth(lay["col_begin"], lay["col_end"], lay["kp"] if world > 1 else k)
    dev_vals = host_vals.to(dev)

    if world == 1:
        out = D.DeviceCommitment(n_log, k, RATE_BITS, CAP_HEIGHT, dev)

        def step():
            D.commit_device(ctx, dev_vals, RATE_BITS, CAP_HEIGHT, out=out)
    else:
        sh = D.ShardedCommitment(ctx, n_log, k, RATE_BITS, CAP_HEIGHT, rank, world, dev)

        def step():
            sh.run(dev_vals)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()

    # ---- timed region (device-resident)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ctx.stage_ms()                      # clear spans
    ctx.set_timing(True)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    t_wall1 = time.time()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    stages = ctx.stage_ms()
    ctx.set_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = cells / (ms_step * 1e-3)

    # ---- end to end through the host-buffer API (pinned host input, H2D + commit + cap D2H per step)
    cap_host = torch.empty((1 << CAP_HEIGHT, 4), dtype=torch.int64).pin_memory()
    if world == 1:
        import ctypes as C
        hctx = z.Context(local_rank)
        lib = hctx._lib

        def e2e_step():
            h = C.c_void_p()
            hctx.check(lib.b200zkp_commit_from_values(hctx._h, C.c_void_p(host_vals.data_ptr()), n_log, k, RATE_BITS,
                                                      CAP_HEIGHT, None, C.byref(h)))
            hctx.check(lib.b200zkp_batch_cap(h, C.c_void_p(cap_host.data_ptr())))
            lib.b200zkp_batch_free(h)
        h2d_bytes = 8 * n * k
    else:
        def e2e_step():
            cap = sh.run_from_host(host_vals, dev_vals)
            cap_host.copy_(cap, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d_bytes = 8 * n * lay["kp"]
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = cells / float(dt.item())
    cap_dev = (out.cap if world == 1 else sh.cap).cpu()
    assert bool((cap_dev == cap_host).all()), "device-resident and end-to-end paths disagree on the cap"
    # the same synthetic input must give the same cap for every N (compare across runs)
    cap_checksum = "%016x" % (int(np.bitwise_xor.reduce(cap_dev.numpy().view(np.uint64).reshape(-1))))

    # ---- strict drop-in: + D2H of coefficients, row-major leaves and digests, overlapped with the hashing (N = 1 only)
    e2e_strict = None
    if world == 1 and not os.environ.get("B200ZKP_SKIP_STRICT"):
        row = k
        outs = [torch.empty(sz, dtype=torch.int64).pin_memory() for sz in (k * n, N * row, 8 * (N - (1 << CAP_HEIGHT)))]

        def strict_step():
            hctx.check(lib.b200zkp_commit_copy_back(hctx._h, C.c_void_p(host_vals.data_ptr()), 0, n_log, k, RATE_BITS, CAP_HEIGHT, None,
                                                    C.c_void_p(outs[0].data_ptr()), C.c_void_p(outs[1].data_ptr()),
                                                    C.c_void_p(outs[2].data_ptr()), C.c_void_p(cap_host.data_ptr()), None))
        strict_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            strict_step()
        torch.cuda.synchronize()
        dt_s = (time.perf_counter() - t0) / max(1, min(args.steps, 3))
        e2e_strict = {"value": cells / dt_s, "unit": UNIT, "ms_per_step": dt_s * 1e3, "h2d_bytes_per_step": 8 * n * k,
                      "d2h_bytes_per_step": sum(o.numel() for o in outs) * 8 + (32 << CAP_HEIGHT),
                      "api": "b200zkp_commit_copy_back: coefficients + row-major leaves + digests + cap to pinned host memory"}
        del outs

    if rank == 0:
        peak, peak_src = measured_peaks()
        N_local = lay["N_local"]
        leaf_ms, leaf_cnt = stages["leaf_hash"]
        leaf_ms_avg = leaf_ms / max(leaf_cnt, 1)
        leaf_bytes = 8 * N_local * k + 32 * N_local          # 64 B/cell read + 32*2^r/k written
        achieved = leaf_bytes / (leaf_ms_avg * 1e-3) / 1e9 if leaf_ms_avg else 0.0
        perms = (N >> 0) * ((k + 7) // 8) + (N - (1 << CAP_HEIGHT))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64 (Goldilocks field, 32-bit IMAD/IADD3 limbs)", "data": "synthetic",
            "config": {"workload": workload_name(n_log, k), "n_log": n_log, "k": k, "rate_bits": RATE_BITS,
                       "cap_height": CAP_HEIGHT, "cells_per_step": cells,
                       "l2": "no flush needed: every step streams 1.13 GB of input and 9 GB of LDE (>> 126 MB L2)",
                       "partition": "single GPU" if world == 1 else
                       f"one commitment over {world} GPUs: column-sharded iNTT, all-gather coeffs, leaf-range LDE+Merkle, all-gather cap"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 32 << CAP_HEIGHT,
                    "api": "b200zkp_commit_from_values + b200zkp_batch_cap (pinned host buffers)" if world == 1 else
                           "ShardedCommitment.run_from_host: pinned column shard H2D (chunked, overlapped with the inverse transform) + commit + cap D2H per rank"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "merkle::leaf_hash_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": leaf_bytes, "avg_launch_ms": leaf_ms_avg,
                         "note": "integer-pipe bound by construction (~1 Poseidon permutation per cell); see int_pipe"},
            "commit_hbm": {"bytes_per_step": algorithmic_bytes(n_log, k),
                           "achieved_gbs": algorithmic_bytes(n_log, k) / (ms_step * 1e-3) / 1e9,
                           "frac_of_peak": algorithmic_bytes(n_log, k) / (ms_step * 1e-3) / 1e9 / (peak * world)},
            "stages_ms_per_step": {s: v[0] / args.steps for s, v in stages.items()},
            "poseidon_perms_per_s": perms / (ms_step * 1e-3),
            "cap_checksum": cap_checksum,
        }
        if e2e_strict:
            line["e2e_strict"] = e2e_strict
        if world == 1 and n_log == N_LOG and k == K:
            # dram__bytes_read.sum + dram__bytes_write.sum of the leaf hash at this size, ncu --set full capture
            # profiles/r1e_leaf_kernel_ncu_details.txt (9.08 GB + 0.27 GB)
            line["roofline"]["traffic"] = 9.36e9
        if world == 1:
            gips = {}
            for kind, name in ((0, "imad_wide"), (1, "iadd3"), (2, "imad"), (3, "imad_wide+lop3"), (4, "lop3"), (5, "imad_hi"), (6, "imad+lop3"), (7, "iadd3_carry_pair"), (8, "imad_wide_noacc")):
                import ctypes as C
                g = C.c_double()
                ctx.check(ctx._lib.b200zkp_int_pipe_bench(ctx._h, kind, 2000, C.byref(g)))
                gips[name] = g.value
            line["int_pipe"] = {"unit": "giga thread-instructions/s (dependent chains, all SMs)", **gips}
            # second roof of SURVEY.md 8d: the leaf hash against the measured issue rate of the integer-multiply (fmaheavy) pipe,
            # the busier of the two integer pipes.  Dynamic instruction mix per permutation of the shipped kernel (ncu source
            # counters, profiles/r1e_leaf_instruction_mix.txt: 17 964 warp-instructions per warp-permutation):
            #   multiplier pipe: 2605 IMAD.WIDE (2 slots each: they issue at half the IMAD rate) + 5574 single-slot
            #                    (2033 IMAD shift-adds of the linear layers, the rest IMAD.IADD / IMAD.X / IMAD.MOV / IMAD.SHL)
            #   ALU pipe:        9217 IADD3 / IADD3.X / LOP3 / SEL / LEA / SHF (one slot each)
            leaf_perms = N_local * ((k + 7) // 8)
            slots_per_perm = 2 * 2605 + 5574
            alu_slots_per_perm = 9217
            instr_per_perm = 17964
            slot_rate = leaf_perms * slots_per_perm / (leaf_ms_avg * 1e-3) if leaf_ms_avg else 0.0
            slot_peak = gips["imad"] * 1e9
            line["roofline"]["int"] = {"bound": "integer multiply pipe (fmaheavy)", "achieved": slot_rate, "peak": slot_peak,
                                       "unit": "IMAD-slot thread-instr/s (IMAD.WIDE = 2 slots)", "frac": slot_rate / slot_peak,
                                       "slots_per_permutation": slots_per_perm, "imad_wide_per_permutation": 2605,
                                       "alu_slots_per_permutation": alu_slots_per_perm,
                                       "alu_frac": (leaf_perms * alu_slots_per_perm / (leaf_ms_avg * 1e-3) / (gips["lop3"] * 1e9)) if leaf_ms_avg else 0.0,
                                       "instructions_per_permutation": instr_per_perm,
                                       "issue_frac": (leaf_perms * instr_per_perm / (leaf_ms_avg * 1e-3) / (2 * gips["imad"] * 1e9)) if leaf_ms_avg else 0.0,
                                       "permutations_per_launch": leaf_perms,
                                       "ncu": {"pipe_fmaheavy_busy": 0.843, "pipe_alu_busy": 0.666, "issue_active": 0.679}}
            if not os.environ.get("B200ZKP_SKIP_CPU"):
                from oracle import oracle as O
                O.build()
                line["cpu_baseline"] = cpu_baseline_sampled()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
