"""Device-resident entry points: torch tensors in HBM in, torch tensors out, no host copies.

torch is used for what the north star allows it for — device memory, streams and torch.distributed (NCCL);
every transform and hash is one of this repo's CUDA kernels behind the C ABI (`b200zkp_dev_*`).

Multi-GPU partitioning (one process per GPU):
  1. values are sharded by COLUMN: rank g inverse-transforms columns [g*kp, (g+1)*kp), kp = ceil(k/G)
  2. one all-gather of the coefficients (NCCL over NVLink; k padded to G*kp with zero columns)
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
    `rank` inverse-transforms, which coset blocks / leaf range / cap entries it owns."""
    if world <= 0 or world & (world - 1) or world > (1 << rate_bits) or world > (1 << cap_height):
        raise ValueError("world size must be a power of two <= 2^rate_bits and <= 2^cap_height")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    N = 1 << (n_log + rate_bits)
    kp = (k + world - 1) // world
    bpr = (1 << rate_bits) // world
    cpr = (1 << cap_height) // world
    return dict(
        kp=kp, col_begin=min(k, rank * kp), col_end=min(k, (rank + 1) * kp),
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


class ShardedCommitment:
    """One rank's share of a commitment partitioned over `world` GPUs (see module docstring)."""

    def __init__(self, ctx: Context, n_log: int, k: int, rate_bits: int, cap_height: int, rank: int, world: int,
                 device, group=None):
        import torch
        lay = shard_layout(n_log, k, rate_bits, cap_height, rank, world)
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.n_log, self.k, self.rate_bits, self.cap_height = n_log, k, rate_bits, cap_height
        n = 1 << n_log
        self.kp = lay["kp"]                                     # columns per rank (padded)
        self.blocks_per_rank = lay["block_end"] - lay["block_begin"]
        self.N_local = lay["N_local"]
        self.cap_height_local = lay["cap_height_local"]
        self.col_begin, self.col_end = lay["col_begin"], lay["col_end"]
        self.coeffs_all = torch.zeros((self.kp * world, n), dtype=torch.int64, device=device)
        self.lde = torch.empty((k, self.N_local), dtype=torch.int64, device=device)
        self.digests = torch.empty((max(2 * (self.N_local - (1 << self.cap_height_local)), 1), 4),
                                   dtype=torch.int64, device=device)
        self.cap_local = torch.empty((1 << self.cap_height_local, 4), dtype=torch.int64, device=device)
        self.cap = torch.empty((1 << cap_height, 4), dtype=torch.int64, device=device)

    @property
    def local_columns(self) -> int:
        return max(self.col_end - self.col_begin, 0)

    def run_from_host(self, host_values, staging, n_chunks: int = 4):
        """End-to-end variant: `host_values` is this rank's pinned (kp, n) column shard, `staging` a (kp, n) device
        tensor.  Column chunks cross PCIe on a side stream while the previous chunk is inverse-transformed."""
        import torch
        ctx, lib = self.ctx, self.ctx._lib
        n = 1 << self.n_log
        kl = self.local_columns
        mine = self.coeffs_all[self.rank * self.kp:(self.rank + 1) * self.kp]
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=staging.device)
        main = torch.cuda.current_stream()
        self._copy_stream.wait_stream(main)          # staging may still be read by the previous step
        per = max(1, (kl + n_chunks - 1) // n_chunks)
        events = []
        for c0 in range(0, kl, per):
            c1 = min(kl, c0 + per)
            with torch.cuda.stream(self._copy_stream):
                staging[c0:c1].copy_(host_values[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            events.append((c0, c1, ev))
        for c0, c1, ev in events:
            main.wait_event(ev)
            # scratch: this chunk's share of the (still unused) LDE buffer
            ctx.check(lib.b200zkp_dev_intt(ctx._h, C.c_void_p(staging.data_ptr() + 8 * n * c0), n,
                                           C.c_void_p(mine.data_ptr() + 8 * n * c0), n,
                                           C.c_void_p(self.lde.data_ptr() + 8 * n * c0), self.n_log, c1 - c0))
        return self._finish()

    def run(self, values_local, is_coeffs: bool = False):
        """values_local: (kp, n) device tensor, this rank's columns (rows past local_columns ignored)."""
        ctx, lib = self.ctx, self.ctx._lib
        n = 1 << self.n_log
        mine = self.coeffs_all[self.rank * self.kp:(self.rank + 1) * self.kp]
        kl = self.local_columns
        if kl:
            if is_coeffs:
                mine[:kl].copy_(values_local[:kl])
            else:
                # the LDE buffer is free until step 3: use it as the transform scratch
                ctx.check(lib.b200zkp_dev_intt(ctx._h, _ptr(values_local), n, _ptr(mine), n, _ptr(self.lde),
                                               self.n_log, kl))
        return self._finish()

    def _finish(self):
        """all-gather of the coefficients, this rank's coset blocks, its cap subtrees, all-gather of the cap."""
        import torch
        import torch.distributed as dist
        ctx, lib = self.ctx, self.ctx._lib
        n = 1 << self.n_log
        mine = self.coeffs_all[self.rank * self.kp:(self.rank + 1) * self.kp]
        b0 = self.rank * self.blocks_per_rank
        b1 = b0 + self.blocks_per_rank
        N_loc = self.N_local

        def lde_cols(c0, c1):
            if c1 > c0:
                ctx.check(lib.b200zkp_dev_lde(ctx._h, C.c_void_p(self.coeffs_all.data_ptr() + 8 * n * c0), n,
                                              C.c_void_p(self.lde.data_ptr() + 8 * N_loc * c0), N_loc, self.n_log,
                                              c1 - c0, self.rate_bits, b0, b1))

        if self.world > 1:
            # the all-gather of the other ranks' coefficients (NCCL, its own stream) overlaps the coset transforms of
            # the columns this rank already holds
            work = dist.all_gather_into_tensor(self.coeffs_all, mine, group=self.group, async_op=True)
            lde_cols(self.col_begin, self.col_end)
            work.wait()
            lde_cols(0, self.col_begin)
            lde_cols(self.col_end, self.k)
        else:
            lde_cols(0, self.k)
        ctx.check(lib.b200zkp_dev_merkle(ctx._h, _ptr(self.lde), 1, N_loc, self.k, N_loc, self.cap_height_local,
                                         _ptr(self.digests), _ptr(self.cap_local)))
        if self.world > 1:
            dist.all_gather_into_tensor(self.cap, self.cap_local, group=self.group)
        else:
            self.cap.copy_(self.cap_local)
        return self.cap

    def rows(self, indices):
        """MerkleTree::get + MerkleTree::prove for GLOBAL leaf indices of the sharded commitment (what the FRI query rounds
        open): the rank that owns a leaf gathers its row and sibling path from its shard, one all-reduce (sum with zeros from
        the other ranks) hands every rank the full answer.  Returns (rows (q, k), siblings (q, log2 N - cap_height, 4))."""
        import torch
        import torch.distributed as dist
        ctx, lib = self.ctx, self.ctx._lib
        dev = self.lde.device
        idx = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
        q = idx.numel()
        depth = log2_strict(self.N_local) - self.cap_height_local
        rows = torch.zeros((q, self.k), dtype=torch.int64, device=dev)
        sib = torch.zeros((q, depth, 4), dtype=torch.int64, device=dev)
        if q == 0:
            return rows, sib
        if int(idx.min()) < 0 or int(idx.max()) >= self.N_local * self.world:
            raise ValueError("leaf index out of range")
        mine = (idx // self.N_local) == self.rank
        local = torch.where(mine, idx % self.N_local, torch.zeros_like(idx)).contiguous()
        ctx.check(lib.b200zkp_dev_gather(ctx._h, _ptr(self.lde), self.N_local, self.k, _ptr(self.digests), self.N_local,
                                         self.cap_height_local, _ptr(local), q, _ptr(rows), _ptr(sib) if depth else None))
        rows *= mine[:, None]
        sib *= mine[:, None, None]
        if self.world > 1:
            dist.all_reduce(rows, group=self.group)
            dist.all_reduce(sib, group=self.group)
        return rows, sib
