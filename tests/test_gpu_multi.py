"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): two NCCL ranks run the sharded commitment of
intmax_zkp_core_b200.device.ShardedCommitment and must reproduce the oracle's cap, leaves and digests."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_log, k, r, h, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import oracle as O
    from intmax_zkp_core_b200 import device as D
    n = 1 << n_log
    v = O.synthetic_values(k, n, seed=3)
    ref = O.commit(v, r, h)
    ctx = D.torch_context(rank)
    lay = D.shard_layout(n_log, k, r, h, rank, world)
    mine = np.zeros((lay["kp"], n), np.uint64)
    mine[:lay["col_end"] - lay["col_begin"]] = v[lay["col_begin"]:lay["col_end"]]
    sh = D.ShardedCommitment(ctx, n_log, k, r, h, rank, world, dev)
    cap = sh.run(torch.from_numpy(mine.view(np.int64)).to(dev))
    torch.cuda.synchronize()
    ok = bool((cap.cpu().numpy().view(np.uint64) == ref["cap"]).all())
    ok &= bool((sh.lde.cpu().numpy().view(np.uint64).T == ref["leaves"][lay["leaf_begin"]:lay["leaf_end"]]).all())
    ok &= bool((sh.coeffs_all.cpu().numpy().view(np.uint64)[:k] == ref["coeffs"]).all())
    sub = 2 * (((n << r) >> h) - 1)
    ok &= bool((sh.digests.cpu().numpy().view(np.uint64)[:(lay["cap_end"] - lay["cap_begin"]) * sub]
                == ref["digests"][lay["cap_begin"] * sub:lay["cap_end"] * sub]).all())
    # (e) step 5: openings of global leaf indices, gathered by the owning rank and shared with one all-reduce
    N = n << r
    idx = [0, 5, N // 2 - 1, N // 2, N - 1, 17 % N]
    rows, sib = sh.rows(idx)
    torch.cuda.synchronize()
    rows, sib = rows.cpu().numpy().view(np.uint64), sib.cpu().numpy().view(np.uint64)
    for j, x in enumerate(idx):
        ok &= bool((rows[j] == ref["leaves"][x]).all())
        ok &= bool((sib[j] == O.merkle_prove(ref["digests"], N, h, x)).all())
        ok &= bool(O.merkle_verify(rows[j], x, sib[j], ref["cap"]))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_log,k,r,h", [(10, 135, 3, 4), (13, 7, 3, 4), (6, 3, 1, 1)])
def test_two_gpu_sharded_commitment(n_log, k, r, h):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, n_log, k, r, h, q)) for rk in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
