"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise) through the C ABI's partitioned commitment
(b200zkp_comm_* / b200zkp_sharded_*, host mirror intmax_zkp_core_b200.device.Comm / ShardedCommitment): cap, leaves, digests,
coefficients and openings must equal the oracle's unsharded commitment, both with one process per GPU
(b200zkp_comm_init_rank) and with one process driving every GPU (b200zkp_comm_init_all)."""
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


def _check_rank(O, sh, i, lay, ref, n_log, k, r, h):
    """cap, leaves, coefficients and digests of local rank i against the oracle's unsharded commitment"""
    import torch
    n = 1 << n_log
    v = sh.local(i)
    with torch.cuda.device(v["cap"].device):
        torch.cuda.synchronize()
    ok = bool((v["cap"].cpu().numpy().view(np.uint64) == ref["cap"]).all())
    ok &= bool((v["lde"].cpu().numpy().view(np.uint64).T == ref["leaves"][lay["leaf_begin"]:lay["leaf_end"]]).all())
    ok &= bool((v["coeffs_all"].cpu().numpy().view(np.uint64)[:k] == ref["coeffs"]).all())
    sub = 2 * (((n << r) >> h) - 1)
    if sub:
        ok &= bool((v["digests"].cpu().numpy().view(np.uint64)[:(lay["cap_end"] - lay["cap_begin"]) * sub]
                    == ref["digests"][lay["cap_begin"] * sub:lay["cap_end"] * sub]).all())
    return ok


def _check_rows(O, sh, ref, n_log, r, h):
    # (e) step 5: openings of GLOBAL leaf indices, answered by the owning rank (b200zkp_sharded_rows)
    N = (1 << n_log) << r
    idx = [0, 5 % N, N // 2 - 1, N // 2, N - 1, 17 % N]
    rows, sib = sh.rows(idx)
    ok = True
    for j, x in enumerate(idx):
        ok &= bool((rows[j] == ref["leaves"][x]).all())
        ok &= bool((sib[j] == O.merkle_prove(ref["digests"], N, h, x)).all())
        ok &= bool(O.merkle_verify(rows[j], x, sib[j], ref["cap"]))
    return ok


def _worker(rank, world, port, n_log, k, r, h, q):
    """one process per GPU: the NCCL id travels over a gloo group, everything else is the C ABI (b200zkp_comm_init_rank)"""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from intmax_zkp_core_b200 import device as D
    n = 1 << n_log
    v = O.synthetic_values(k, n, seed=3)
    ref = O.commit(v, r, h)
    ctx = D.torch_context(rank)
    comm = D.Comm.from_torch_distributed(ctx)
    lay = D.shard_layout(n_log, k, r, h, rank, world)
    mine = np.ascontiguousarray(v[lay["cols"]])
    sh = D.ShardedCommitment(comm, n_log, k, r, h)
    # device-resident input (asynchronous), twice: the second commit reuses every buffer
    dv = torch.from_numpy(mine.view(np.int64)).to(dev) if mine.size else None
    sh.run(dv)
    sh.run(dv)
    ok = _check_rank(O, sh, 0, lay, ref, n_log, k, r, h)
    # host input, cap written to the caller's buffer before the call returns
    cap_host = torch.zeros((1 << h, 4), dtype=torch.int64).pin_memory()
    hv = torch.from_numpy(mine.view(np.int64)).pin_memory() if mine.size else None
    sh.run_from_host(hv, cap_out=cap_host)
    ok &= bool((cap_host.numpy().view(np.uint64) == ref["cap"]).all())
    ok &= _check_rank(O, sh, 0, lay, ref, n_log, k, r, h)
    ok &= _check_rows(O, sh, ref, n_log, r, h)
    # from_coeffs form
    cf = np.ascontiguousarray(ref["coeffs"][lay["cols"]])
    sh.run(torch.from_numpy(cf.view(np.int64)).to(dev) if cf.size else None, is_coeffs=True)
    ok &= _check_rank(O, sh, 0, lay, ref, n_log, k, r, h)
    # other data through the same buffers and exchange windows (a stale read of a peer's window would show here), in both
    # forms of the exchange: peer memory inside the transform (CUDA IPC between these processes) and NCCL point-to-point
    peer = comm.peer_exchange
    v2 = O.synthetic_values(k, n, seed=4)
    ref2 = O.commit(v2, r, h)
    mine2 = np.ascontiguousarray(v2[lay["cols"]])
    dv2 = torch.from_numpy(mine2.view(np.int64)).to(dev) if mine2.size else None
    hv2 = torch.from_numpy(mine2.view(np.int64)).pin_memory() if mine2.size else None
    for form in ((True, False) if peer else (False,)):
        comm.set_peer_exchange(form)
        ok &= comm.peer_exchange == form
        sh.run(dv2)
        ok &= _check_rank(O, sh, 0, lay, ref2, n_log, k, r, h)
        sh.run(dv)
        ok &= _check_rank(O, sh, 0, lay, ref, n_log, k, r, h)
        cap_host.zero_()
        sh.run_from_host(hv2, cap_out=cap_host)
        ok &= bool((cap_host.numpy().view(np.uint64) == ref2["cap"]).all())
        ok &= _check_rank(O, sh, 0, lay, ref2, n_log, k, r, h)
    sh.close()
    comm.close()
    q.put((rank, ok, peer))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_log,k,r,h", [(10, 135, 3, 4), (13, 7, 3, 4), (6, 3, 1, 1), (14, 21, 3, 4), (12, 5, 1, 2)])
def test_two_gpu_sharded_commitment(n_log, k, r, h, monkeypatch):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("B200ZKP_PEER_CHUNK_BYTES", str(3 * 8 << n_log))     # host inputs: chunks of 3 columns
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(rk, world, port, n_log, k, r, h, q)) for rk in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r_[:2] for r_ in res) == [(0, True), (1, True)]
    print("peer exchange between processes:", [r_[2] for r_ in res])


@pytest.mark.parametrize("n_log,k,r,h", [(10, 135, 3, 4), (12, 7, 2, 3), (4, 3, 1, 1), (13, 19, 3, 4)])
def test_one_process_drives_all_gpus(oracle, n_log, k, r, h, monkeypatch):
    """b200zkp_comm_init_all: the deployment a single-process (rayon) prove() needs"""
    import torch
    import intmax_zkp_core_b200 as z
    from intmax_zkp_core_b200 import device as D
    O = oracle
    G = 2
    while G * 2 <= min(torch.cuda.device_count(), 1 << r, 1 << h):
        G *= 2
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("B200ZKP_PEER_CHUNK_BYTES", str(2 * 8 << n_log))     # host inputs: chunks of 2 columns
    ctxs = [z.Context(g) for g in range(G)]
    comm = D.Comm.init_all(ctxs)
    assert (comm.world, comm.n_local, comm.rank0) == (G, G, 0)
    n = 1 << n_log
    v = O.synthetic_values(k, n, seed=5)
    ref = O.commit(v, r, h)
    sh = D.ShardedCommitment(comm, n_log, k, r, h)
    lays = sh.layouts
    host = [torch.from_numpy(np.ascontiguousarray(v[l["cols"]]).view(np.int64)).pin_memory() if l["n_cols"] else None for l in lays]
    cap_host = torch.zeros((1 << h, 4), dtype=torch.int64).pin_memory()
    assert comm.peer_exchange, "peer access between the GPUs of one box"
    v2 = O.synthetic_values(k, n, seed=6)
    ref2 = O.commit(v2, r, h)
    host2 = [torch.from_numpy(np.ascontiguousarray(v2[l["cols"]]).view(np.int64)).pin_memory() if l["n_cols"] else None for l in lays]
    for groups in (None, 2, 0, 1):          # None: peer-memory exchange; else the NCCL exchange with that group size
        comm.set_peer_exchange(groups is None)
        assert comm.peer_exchange == (groups is None)
        if groups is not None:
            comm.set_exchange_group(groups)
        for hh, rr in ((host, ref), (host2, ref2), (host, ref)):
            cap_host.zero_()
            sh.run_from_host(hh, cap_out=cap_host)
            assert (cap_host.numpy().view(np.uint64) == rr["cap"]).all()
            for i, lay in enumerate(lays):
                assert _check_rank(O, sh, i, lay, rr, n_log, k, r, h), (groups, i)
    comm.set_peer_exchange(True)
    assert _check_rows(O, sh, ref, n_log, r, h)
    dv = [h_.to(torch.device("cuda", i)) if h_ is not None else None for i, h_ in enumerate(host)]
    sh.run(dv)
    sh.synchronize()
    for i, lay in enumerate(lays):
        assert _check_rank(O, sh, i, lay, ref, n_log, k, r, h)
    sh.close()
    # the one-call form on the full host matrix
    import ctypes as C
    hsh, cap2 = C.c_void_p(), np.zeros((1 << h, 4), np.uint64)
    vv = np.ascontiguousarray(v)
    comm.check(comm._lib.b200zkp_sharded_commit_from_values(comm._h, vv.ctypes.data_as(C.c_void_p), n_log, k, r, h,
                                                            cap2.ctypes.data_as(C.c_void_p), C.byref(hsh)))
    assert (cap2 == ref["cap"]).all()
    comm._lib.b200zkp_sharded_free(hsh)
    # plonky2-style argument errors
    with pytest.raises(ValueError):
        D.ShardedCommitment(comm, 3, 5, 0, 4)          # more ranks than coset blocks
    comm.close()
    for c in ctxs:
        c.close()
