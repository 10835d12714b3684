"""Device-resident entry points: torch tensors in HBM in, torch tensors out, no host copies.

torch is used for what the north star allows it for — device memory and streams; every transform and hash is one of this
repo's CUDA kernels behind the C ABI (`b200zkp_dev_*`), and the multi-GPU exchange is NCCL called from the C library
(`b200zkp_comm_*`, `b200zkp_sharded_*`; csrc/sharded.inl), not torch.distributed.

Multi-GPU partitioning (one process per GPU):
  1. values are sharded by COLUMN: rank g inverse-transforms columns [g*kp, (g+1)*kp), kp = ceil(k/G)
  2. the coefficient shards are exchanged in NCCL point-to-point groups (k padded to G*kp with zero columns); the coset
     transforms of the shards already received overlap the rest of the exchange
  3. the LDE is sharded by LEAF RANGE: rank g owns leaves [g*N/G, (g+1)*N/G) = 2^r/G whole cosets of the
     same coefficients, so the coset NTTs, the leaf hashing and the 2^h/G cap subtrees are all local
  4. one all-gather of the 2^h x 32 B cap digests.
Requires G to be a power of two with G <= 2^rate_bits and G <= 2^cap_height (8 GPUs at r = 3, h = 4).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _lib
from .plonky2 import Context, log2_strict


def _ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else None


def torch_context(device_index: int) -> Context:
    """A Context that enqueues on torch's current stream of `device_index` (so torch.cuda.Event sees it)."""
    import torch
    with torch.cuda.device(device_index):
        stream = torch.cuda.current_stream().cuda_stream
    # stream 0 (legacy default) is a valid shared stream, but ctx_create treats NULL as "make my own":
    # torch's default stream handle is 0, so create a dedicated torch stream instead and make it current.
    if stream == 0:
        s = torch.cuda.Stream(device=device_index)
        torch.cuda.set_stream(s)
        stream = s.cuda_stream
        ctx = Context(device_index, stream)
        ctx._torch_stream = s
        return ctx
    return Context(device_index, stream)


def shard_layout(n_log: int, k: int, rate_bits: int, cap_height: int, rank: int, world: int) -> dict:
    """Pure partition arithmetic of the multi-GPU commitment (no torch, no device): which columns rank
    `rank` inverse-transforms, which coset blocks / leaf range / cap entries it owns.
    Columns are dealt in groups of L = max(8, world) — a whole number of sponge chunks, so a group can be hashed as soon as it is
    complete — w = L / world consecutive columns of every group per rank: `cols` lists the rank's columns in increasing order,
    and that is the order of the rows of the shard it hands to ShardedCommitment.run / run_from_host."""
    if world <= 0 or world & (world - 1) or world > (1 << rate_bits) or world > (1 << cap_height):
        raise ValueError("world size must be a power of two <= 2^rate_bits and <= 2^cap_height")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    N = 1 << (n_log + rate_bits)
    L = max(8, world)
    w = L // world
    kp = -(-k // L) * w
    cols = [c for c in range(k) if (c % L) // w == rank]
    bpr = (1 << rate_bits) // world
    cpr = (1 << cap_height) // world
    return dict(
        kp=kp, group=L, per_group=w, cols=cols, n_cols=len(cols),
        block_begin=rank * bpr, block_end=(rank + 1) * bpr,
        N_local=N // world, leaf_begin=rank * (N // world), leaf_end=(rank + 1) * (N // world),
        cap_height_local=cap_height - log2_strict(world), cap_begin=rank * cpr, cap_end=(rank + 1) * cpr,
    )


class DeviceCommitment:
    """Buffers of one commitment, all in HBM (torch int64 tensors holding uint64 bit patterns)."""

    def __init__(self, n_log: int, k: int, rate_bits: int, cap_height: int, device, salt: bool = False):
        import torch
        n, N = 1 << n_log, 1 << (n_log + rate_bits)
        row = k + (4 if salt else 0)
        self.n_log, self.k, self.rate_bits, self.cap_height, self.row = n_log, k, rate_bits, cap_height, row
        self.coeffs = torch.empty((k, n), dtype=torch.int64, device=device)
        self.lde = torch.empty((row, N), dtype=torch.int64, device=device)
        self.digests = torch.empty((max(2 * (N - (1 << cap_height)), 1), 4), dtype=torch.int64, device=device)
        self.cap = torch.empty((1 << cap_height, 4), dtype=torch.int64, device=device)


def commit_device(ctx: Context, inp, rate_bits: int, cap_height: int, out: Optional[DeviceCommitment] = None,
                  is_coeffs: bool = False, salt=None) -> DeviceCommitment:
    """PolynomialBatch::from_values/from_coeffs on device tensors; asynchronous on the ctx stream."""
    k, n = inp.shape
    n_log = log2_strict(n)
    assert inp.is_cuda and inp.is_contiguous() and inp.element_size() == 8
    if out is None:
        out = DeviceCommitment(n_log, k, rate_bits, cap_height, inp.device, salt is not None)
    ctx.check(ctx._lib.b200zkp_dev_commit(ctx._h, _ptr(inp), int(is_coeffs), n_log, k, rate_bits, cap_height,
                                          _ptr(salt), _ptr(out.coeffs), _ptr(out.lde), _ptr(out.digests),
                                          _ptr(out.cap)))
    return out


def _dev_view(ptr: int, shape, device_index: int):
    """torch int64 view of `shape` words of device memory owned by the C library (valid while its handle lives)"""
    import torch

    class _View:
        pass
    v = _View()
    v.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(v, device=torch.device("cuda", device_index))


class Comm:
    """b200zkp_comm: the ranks of one partitioned commitment.  NCCL is driven from the C library (csrc/sharded.inl);
    torch.distributed is at most the out-of-band channel that hands rank 0's NCCL id to the other processes."""

    def __init__(self, ctxs, handle):
        self.ctxs, self._h = list(ctxs), handle
        self._lib = self.ctxs[0]._lib
        shape = (C.c_int32 * 3)()
        self._lib.b200zkp_comm_shape(self._h, shape)
        self.world, self.n_local, self.rank0 = int(shape[0]), int(shape[1]), int(shape[2])

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _lib.lib().b200zkp_comm_unique_id(buf)
        if rc != 0:
            raise _lib.B200ZkpError(rc, "b200zkp_comm_unique_id failed (is libnccl.so.2 available?)")
        return bytes(buf)

    @classmethod
    def init_rank(cls, ctx: Context, unique_id: bytes, rank: int, world: int) -> "Comm":
        """one process per GPU: every process calls this with the same id (b200zkp_comm_init_rank)"""
        h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        ctx.check(ctx._lib.b200zkp_comm_init_rank(ctx._h, buf, rank, world, C.byref(h)))
        return cls([ctx], h)

    @classmethod
    def from_torch_distributed(cls, ctx: Context, group=None) -> "Comm":
        """one process per GPU under torchrun: rank 0's id travels through the (gloo or nccl) process group"""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls.init_rank(ctx, box[0], rank, world)

    @classmethod
    def init_all(cls, ctxs) -> "Comm":
        """ONE process drives len(ctxs) GPUs (b200zkp_comm_init_all): ctxs[i] is rank i"""
        ctxs = list(ctxs)
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        h = C.c_void_p()
        ctxs[0].check(ctxs[0]._lib.b200zkp_comm_init_all(arr, len(ctxs), C.byref(h)))
        return cls(ctxs, h)

    def check(self, rc: int):
        if rc != 0:
            msg = self._lib.b200zkp_comm_last_error(self._h).decode() or self._lib.b200zkp_last_error(self.ctxs[0]._h).decode()
            raise _lib.B200ZkpError(rc, msg)

    def set_exchange_group(self, peers_per_group: int):
        self.check(self._lib.b200zkp_comm_set_exchange_group(self._h, peers_per_group))

    def set_peer_exchange(self, enabled: bool):
        """peer-memory form of the coefficient exchange (default where available) or the NCCL form; collective"""
        self.check(self._lib.b200zkp_comm_set_peer_exchange(self._h, int(bool(enabled))))

    @property
    def peer_exchange(self) -> bool:
        """True when commits gather the coefficients through peer memory inside the transform (b200zkp_comm_peer_exchange)"""
        return bool(self._lib.b200zkp_comm_peer_exchange(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200zkp_comm_destroy(self._h)
            self._h = None


class ShardedCommitment:
    """One commitment partitioned over the ranks of a Comm (b200zkp_sharded; see the module docstring).  With one process
    per GPU this object holds that process's rank; with Comm.init_all it holds every rank (index them with `local`)."""

    def __init__(self, comm: Comm, n_log: int, k: int, rate_bits: int, cap_height: int):
        self.comm, self._lib = comm, comm._lib
        self.n_log, self.k, self.rate_bits, self.cap_height = n_log, k, rate_bits, cap_height
        shard_layout(n_log, k, rate_bits, cap_height, 0, comm.world)          # plonky2-style argument errors as ValueError
        h = C.c_void_p()
        comm.check(self._lib.b200zkp_sharded_create(comm._h, n_log, k, rate_bits, cap_height, C.byref(h)))
        self._h = h
        self.layouts = [shard_layout(n_log, k, rate_bits, cap_height, comm.rank0 + i, comm.world) for i in range(comm.n_local)]
        lay = self.layouts[0]
        self.kp, self.N_local, self.cap_height_local = lay["kp"], lay["N_local"], lay["cap_height_local"]
        self.world = comm.world
        n = 1 << n_log
        self._views = []
        for i in range(comm.n_local):
            p = [C.c_void_p() for _ in range(4)]
            comm.check(self._lib.b200zkp_sharded_device_ptrs(self._h, i, *[C.byref(x) for x in p]))
            dev = comm.ctxs[i].device
            nd = max(2 * (self.N_local - (1 << self.cap_height_local)), 1)
            self._views.append(dict(
                coeffs_all=_dev_view(p[0].value, (self.kp * comm.world, n), dev),
                lde=_dev_view(p[1].value, (k, self.N_local), dev),
                digests=_dev_view(p[2].value, (nd, 4), dev) if p[2].value else None,
                cap=_dev_view(p[3].value, (1 << cap_height, 4), dev)))

    def local(self, i: int = 0) -> dict:
        """device views of local rank i: coeffs_all (G*kp, n), lde (k, N_local), digests, cap"""
        return self._views[i]

    coeffs_all = property(lambda self: self._views[0]["coeffs_all"])
    lde = property(lambda self: self._views[0]["lde"])
    digests = property(lambda self: self._views[0]["digests"])
    cap = property(lambda self: self._views[0]["cap"])

    def _ptrs(self, tensors):
        if not isinstance(tensors, (list, tuple)):
            tensors = [tensors]
        if len(tensors) != self.comm.n_local:
            raise ValueError("one input per local rank")
        n = 1 << self.n_log
        for t, lay in zip(tensors, self.layouts):
            kl = lay["n_cols"]
            if t is None:
                if kl:
                    raise ValueError("missing input shard")
                continue
            if t.element_size() != 8 or not t.is_contiguous() or t.shape[-1] != n or t.shape[0] < kl:
                raise ValueError("input shard must be a contiguous (>= local columns, n) 64-bit tensor")
        return (C.c_void_p * len(tensors))(*[(t.data_ptr() if t is not None else None) for t in tensors]), tensors

    def run(self, values_local, is_coeffs: bool = False):
        """values_local: this rank's (>= local columns, n) DEVICE tensor (a list, one per local rank, after init_all).
        Asynchronous on the ctx streams; returns the device view of the cap."""
        arr, keep = self._ptrs(values_local)
        self.comm.check(self._lib.b200zkp_sharded_commit(self._h, arr, 1, int(is_coeffs), None))
        return self.cap

    def run_from_host(self, host_values, is_coeffs: bool = False, cap_out=None):
        """host_values: pinned (>= local columns, n) HOST tensor(s); the upload is chunked and overlaps the inverse
        transforms.  cap_out: optional pinned (2^h, 4) host tensor — then the call returns when the cap is in it."""
        arr, keep = self._ptrs(host_values)
        self.comm.check(self._lib.b200zkp_sharded_commit(self._h, arr, 0, int(is_coeffs),
                                                         C.c_void_p(cap_out.data_ptr()) if cap_out is not None else None))
        return self.cap

    def synchronize(self):
        self.comm.check(self._lib.b200zkp_sharded_synchronize(self._h))

    def rows(self, indices):
        """MerkleTree::get + MerkleTree::prove for GLOBAL leaf indices (b200zkp_sharded_rows; collective).
        Returns numpy (rows (q, k), siblings (q, log2 N - cap_height, 4))."""
        import numpy as np
        idx = np.ascontiguousarray(np.asarray(list(indices), dtype=np.uint64))
        depth = log2_strict(self.N_local) - self.cap_height_local
        rows = np.zeros((idx.size, self.k), np.uint64)
        sib = np.zeros((idx.size, depth, 4), np.uint64)
        if idx.size and int(idx.max()) >= self.N_local * self.world:
            raise ValueError("leaf index out of range")
        self.comm.check(self._lib.b200zkp_sharded_rows(self._h, idx.ctypes.data_as(C.c_void_p), idx.size,
                                                       rows.ctypes.data_as(C.c_void_p),
                                                       sib.ctypes.data_as(C.c_void_p) if depth else None))
        return rows, sib

    def close(self):
        if getattr(self, "_h", None):
            self._views = []
            self._lib.b200zkp_sharded_free(self._h)
            self._h = None
