"""N > 1 host logic on CPU: two gloo ranks walk the multi-GPU partition of intmax_zkp_core_b200.device
(column-sharded iNTT -> all-gather coefficients -> leaf-range-sharded LDE + local cap subtrees -> all-gather cap)
with the oracle standing in for the kernels, and must reproduce the unsharded oracle commitment."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_log, k, r, h, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from intmax_zkp_core_b200.device import shard_layout
    from helpers import bitrev_perm
    n, N = 1 << n_log, 1 << (n_log + r)
    lay = shard_layout(n_log, k, r, h, rank, world)
    values = O.synthetic_values(k, n, seed=2)
    # 1. column shard of the inverse transform (zero-padded to kp columns): the rank's columns, w of every group of L
    mine = np.zeros((lay["kp"], n), np.uint64)
    for i, c in enumerate(lay["cols"]):
        mine[i] = O.ifft(values[c])
    # 2. all-gather of coefficients; column c is local column (c // L) * w + c % w of rank (c % L) // w
    gathered = torch.zeros((world * lay["kp"], n), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(mine.view(np.int64)))
    shards = gathered.numpy().view(np.uint64).reshape(world, lay["kp"], n)
    L, w = lay["group"], lay["per_group"]
    coeffs = np.stack([shards[(c % L) // w][(c // L) * w + c % w] for c in range(k)])
    # 3. this rank's coset blocks: block b = leaves [b*n, (b+1)*n) = size-n coset NTT in bit-reversed order
    perm_n = bitrev_perm(n_log)
    g2 = 1753635133440165772
    wN = pow(g2, 1 << (32 - n_log - r), O.P)
    local = np.zeros((lay["N_local"], k), np.uint64)
    for bi, b in enumerate(range(lay["block_begin"], lay["block_end"])):
        t = int(format(b, f"0{r}b")[::-1], 2) if r else 0
        s = 7 * pow(wN, t, O.P) % O.P
        for c in range(k):
            scaled = np.array([int(coeffs[c][j]) * pow(s, j, O.P) % O.P for j in range(n)], dtype=np.uint64)
            local[bi * n:(bi + 1) * n, c] = O.fft(scaled)[perm_n]
    dig, cap_local = O.merkle_new(local, lay["cap_height_local"])
    # 4. all-gather of the cap
    cap = torch.zeros((1 << h, 4), dtype=torch.int64)
    dist.all_gather_into_tensor(cap, torch.from_numpy(cap_local.view(np.int64)))
    full = O.commit(values, r, h)
    ok = bool((cap.numpy().view(np.uint64) == full["cap"]).all())
    ok &= bool((local == full["leaves"][lay["leaf_begin"]:lay["leaf_end"]]).all())
    ok &= bool((coeffs == full["coeffs"]).all())
    sub = 2 * ((N >> h) - 1)
    ok &= bool((dig == full["digests"][lay["cap_begin"] * sub:lay["cap_end"] * sub]).all())
    # 5. openings of global leaf indices: the owner answers from its shard, a sum with zeros shares the answer
    depth = (n_log + r) - h
    idx = np.array([0, N // 2 - 1, N // 2, N - 1, 3 % N], dtype=np.int64)
    rows = np.zeros((idx.size, k), np.uint64)
    sib = np.zeros((idx.size, depth, 4), np.uint64)
    for j, x in enumerate(idx):
        if x // lay["N_local"] == rank:
            xl = int(x % lay["N_local"])
            rows[j] = local[xl]
            sib[j] = O.merkle_prove(dig, lay["N_local"], lay["cap_height_local"], xl)
    t_rows, t_sib = torch.from_numpy(rows.view(np.int64)), torch.from_numpy(sib.view(np.int64))
    dist.all_reduce(t_rows)
    dist.all_reduce(t_sib)
    for j, x in enumerate(idx):
        ok &= bool((t_rows.numpy().view(np.uint64)[j] == full["leaves"][x]).all())
        ok &= bool(O.merkle_verify(t_rows.numpy().view(np.uint64)[j], int(x), t_sib.numpy().view(np.uint64)[j], full["cap"]))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_log,k,r,h", [(4, 5, 3, 4), (3, 7, 1, 1), (3, 19, 1, 1)])
def test_two_rank_partition_reproduces_commitment(n_log, k, r, h):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, n_log, k, r, h, q)) for rk in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_layout_covers_everything():
    from intmax_zkp_core_b200.device import shard_layout
    for world in (1, 2, 4, 8):
        lays = [shard_layout(20, 135, 3, 4, g, world) for g in range(world)]
        assert sorted(c for l in lays for c in l["cols"]) == list(range(135))
        assert all(l["n_cols"] <= l["kp"] and l["group"] % 8 == 0 for l in lays)
        # every group of `group` columns holds `per_group` consecutive columns of every rank
        assert all(l["cols"][:l["per_group"]] == list(range(g * l["per_group"], (g + 1) * l["per_group"])) for g, l in enumerate(lays))
        assert sum(l["N_local"] for l in lays) == 1 << 23
        assert [l["block_begin"] for l in lays] == [g * (8 // world) for g in range(world)]
        assert all(l["cap_end"] - l["cap_begin"] == 16 // world for l in lays)
    with pytest.raises(ValueError):
        shard_layout(20, 135, 3, 2, 0, 8)      # more ranks than cap subtrees
    with pytest.raises(ValueError):
        shard_layout(20, 135, 3, 4, 0, 3)
