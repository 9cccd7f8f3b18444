"""Peer memory between the ranks of one box (one process per GPU): a region of every rank's
HBM mapped into every other rank's address space, so kernels dereference a peer GPU's HBM
directly over NVLink / NVSwitch.

Two ways to get the mapping (`GS_PEER_BACKEND`):
  * `symm` (default): torch's symmetric-memory allocator (CUDA VMM: cuMemCreate + shareable handle
    + cuMemMap with 2 MB pages).  Measured on 2 x B200: random 256-byte row gathers from a 3.2-6.4 GB
    peer region run at 750 GB/s, the peer-copy ceiling.
  * `ipc`: plain cudaMalloc regions of the native library exported with legacy CUDA IPC
    (gs_peer_alloc / gs_peer_export / gs_peer_open).  Fine for the small exchange buffers, but the
    importing side maps small pages: the same gathers fall to 28 GB/s once the region exceeds the
    TLB reach (1 GB: 741 GB/s, 3.2 GB: 28 GB/s).

Two users:
  * `DpExchange`    -- the receive slots + flags of the fused all-reduce/clip/SGD kernel
                       (gs_dp_allreduce_clip_sgd), the path's one exchange step;
  * `ShardedTable`  -- a row-partitioned bf16 feature table (BASELINE.json configs[4]); the
                       aggregation kernel reads remote rows through the mapped base pointers.

torch.distributed is plumbing here: it only carries the 64-byte IPC handles between ranks.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import native
from .native import check

_TYPESTR = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1", torch.int16: "<i2", torch.int64: "<i8"}


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view memory it did not allocate."""

    def __init__(self, ptr: int, shape, dtype: torch.dtype, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": _TYPESTR[dtype],
                                         "data": (int(ptr), False), "version": 3, "strides": None}


def as_tensor(ptr: int, shape, dtype: torch.dtype, device, owner=None) -> torch.Tensor:
    """A torch view of `numel(shape)` elements at device address `ptr` (bf16 via an int16 view)."""
    if dtype == torch.bfloat16:
        return as_tensor(ptr, shape, torch.int16, device, owner).view(torch.bfloat16)
    return torch.as_tensor(_RawCuda(ptr, shape, dtype, owner), device=device)


class PeerRegion:
    """`nbytes` of zeroed device memory on this rank plus mappings of the same region of every
    other rank of `group`.  ptrs[r] is the address of rank r's region as seen from this rank."""

    def __init__(self, nbytes: int, device, group=None, world: Optional[int] = None, rank: Optional[int] = None,
                 backend: Optional[str] = None):
        import os
        import torch.distributed as dist
        self.lib = native.load()
        self.device = torch.device(device)
        self.nbytes = int(nbytes)
        self.group = group
        distributed = dist.is_available() and dist.is_initialized()
        self.world = int(world if world is not None else (dist.get_world_size(group) if distributed else 1))
        self.rank = int(rank if rank is not None else (dist.get_rank(group) if distributed else 0))
        self.backend = backend or os.environ.get("GS_PEER_BACKEND", "symm")
        if self.backend not in ("symm", "ipc"):
            raise ValueError("peer backend must be 'symm' or 'ipc'")
        self._symm = None
        self._opened: List[int] = []
        if self.backend == "symm":
            with torch.cuda.device(self.device):
                if self.world > 1:
                    import torch.distributed._symmetric_memory as symm
                    buf = symm.empty((self.nbytes,), dtype=torch.uint8, device=self.device)
                    hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
                    buf.zero_()
                    torch.cuda.synchronize(self.device)
                    dist.barrier(group=group)
                    self._symm = (buf, hdl)
                    self.ptrs = [int(p) for p in hdl.buffer_ptrs]
                else:
                    buf = torch.zeros((self.nbytes,), dtype=torch.uint8, device=self.device)
                    self._symm = (buf, None)
                    self.ptrs = [buf.data_ptr()]
                self.local = self.ptrs[self.rank]
            return
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
            p = ctypes.c_void_p()
            check(self.lib.gs_peer_alloc(self.nbytes, ctypes.byref(p)), "gs_peer_alloc")
            self.local = int(p.value)
            self.ptrs: List[int] = [0] * self.world
            self.ptrs[self.rank] = self.local
            self._opened: List[int] = []
            if self.world > 1:
                handle = ctypes.create_string_buffer(64)
                check(self.lib.gs_peer_export(ctypes.c_void_p(self.local), handle), "gs_peer_export")
                handles: List[Optional[bytes]] = [None] * self.world
                dist.all_gather_object(handles, handle.raw, group=group)
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    q = ctypes.c_void_p()
                    check(self.lib.gs_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)), f"gs_peer_open(rank {r})")
                    self.ptrs[r] = int(q.value)
                    self._opened.append(int(q.value))
                dist.barrier(group=group)

    def tensor(self, rank: int, offset_bytes: int, shape, dtype: torch.dtype) -> torch.Tensor:
        """A torch view into THIS rank's region (peer regions are for kernels of the library only)."""
        if rank != self.rank:
            raise ValueError("only the local region can be viewed as a tensor")
        if self._symm is not None:
            n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            return self._symm[0][int(offset_bytes):int(offset_bytes) + n].view(dtype).view(*shape)
        return as_tensor(self.ptrs[rank] + int(offset_bytes), shape, dtype, self.device, owner=self)

    def close(self):
        """Unmap the peers' regions and free the local one.  Collective when world > 1: no rank
        may free its region while another still has kernels reading it."""
        if self.local == 0:
            return
        import torch.distributed as dist
        if self._symm is not None:
            torch.cuda.synchronize(self.device)
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            self._symm, self.local = None, 0
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            for q in self._opened:
                self.lib.gs_peer_close(ctypes.c_void_p(q))
            self._opened = []
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self.group)
            self.lib.gs_peer_free(ctypes.c_void_p(self.local))
            self.local = 0


class DpExchange:
    """State of the fused data-parallel update for one flat gradient buffer (trainer.py)."""

    GROUPS = 4

    def __init__(self, flat_grad: torch.Tensor, params: Sequence[torch.Tensor], offsets: Sequence[int],
                 groups: Sequence[int], *, world: int = 1, rank: int = 0, group=None, timeout_s: float = 10.0,
                 region_ptrs: Optional[Sequence[int]] = None, params_lo: Optional[Sequence[Optional[torch.Tensor]]] = None,
                 extras: Optional[Sequence[Optional[torch.Tensor]]] = None):
        lib = native.load()
        self.lib, self.flat, self.world, self.rank = lib, flat_grad, int(world), int(rank)
        native.require_cuda(flat_grad, "flat gradient")
        n = len(params)
        self.n_total = int(flat_grad.numel())
        self._params = list(params)                               # keep the storages alive
        self._p = (ctypes.c_void_p * n)(*[p.data_ptr() for p in params])
        self._o = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
        self._n = (ctypes.c_int64 * n)(*[int(p.numel()) for p in params])
        self._g = (ctypes.c_int32 * n)(*[int(g) for g in groups])
        # low halves p - trunc_tf32(p) of selected parameters, rewritten by the update (K4's pre-split weight operand)
        self._params_lo = list(params_lo) if params_lo is not None else None
        self._plo = None
        if self._params_lo is not None:
            self._plo = (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in self._params_lo])
        # extras[k]: [replicas, stride] more partial gradients of tensor k (zeroed; folded in and cleared by every update)
        self._extras = list(extras) if extras is not None else None
        self._ex = self._exn = self._exs = None
        if self._extras is not None:
            self._ex = (ctypes.c_void_p * n)(*[(t.data_ptr() if t is not None else None) for t in self._extras])
            self._exn = (ctypes.c_int32 * n)(*[(int(t.shape[0]) if t is not None else 0) for t in self._extras])
            self._exs = (ctypes.c_int64 * n)(*[(int(t.stride(0)) if t is not None else 0) for t in self._extras])
        self.num_segs = n
        self.state = torch.zeros((int(lib.gs_dp_state_bytes()),), dtype=torch.uint8, device=flat_grad.device)
        self.timeout_ns = int(timeout_s * 1e9)
        self.region: Optional[PeerRegion] = None
        self._regions = None
        if self.world > 1 and region_ptrs is not None:     # caller-provided exchange regions (single-GPU tests)
            self._regions = (ctypes.c_void_p * self.world)(*[int(x) for x in region_ptrs])
        elif self.world > 1:
            nbytes = int(lib.gs_dp_region_bytes(self.n_total, self.world))
            self.region = PeerRegion(nbytes, flat_grad.device, group=group, world=self.world, rank=self.rank)
            self._regions = (ctypes.c_void_p * self.world)(*self.region.ptrs)
            # the receive slots start out EMPTY (every word 0xffffffff): the exchange has no flags, data words are their
            # own arrival signal (csrc/dp_update.cu); no rank may push before every rank has filled its region
            recv_off = int(lib.gs_dp_region_recv_offset())
            self.region.tensor(self.rank, recv_off, (nbytes - recv_off,), torch.uint8).fill_(0xFF)
            torch.cuda.synchronize(flat_grad.device)
            import torch.distributed as dist
            dist.barrier(group=group)

    def update(self, max_norm: float, lr: float, step_counter: Optional[torch.Tensor] = None):
        """all-reduce(mean) -> clip per group -> SGD -> zero gradients (-> step_counter += 1); one launch
        on the current stream."""
        check(self.lib.gs_dp_allreduce_clip_sgd(self.flat.data_ptr(), self.n_total, self._regions, self.rank, self.world,
                                                self._p, self._o, self._n, self._g, self.num_segs, float(max_norm),
                                                float(lr), self.state.data_ptr(), self.timeout_ns,
                                                step_counter.data_ptr() if step_counter is not None else None,
                                                self._plo, self._ex, self._exn, self._exs, native.stream()),
              "gs_dp_allreduce_clip_sgd")

    def status(self):
        """(epoch, status, norms[4]) after synchronising the current stream; raises on a timed-out exchange."""
        e, s = ctypes.c_uint32(), ctypes.c_uint32()
        norms = (ctypes.c_float * self.GROUPS)()
        check(self.lib.gs_dp_status(self.state.data_ptr(), ctypes.byref(e), ctypes.byref(s), norms, native.stream()),
              "gs_dp_status")
        if s.value != 0:
            raise RuntimeError(f"data-parallel exchange timed out (status {s.value}, epoch {e.value}): a peer rank did not "
                               f"arrive within {self.timeout_ns / 1e9:.1f}s")
        return int(e.value), int(s.value), [float(x) for x in norms]

    def close(self):
        if self.region is not None:
            self.region.close()
            self.region = None


class ShardedTable:
    """Row-partitioned bf16 feature table: node v lives in shard v // rows_per_shard at row
    v % rows_per_shard.  Each rank owns one shard in its own HBM; the others are peer mappings.

    `GraphSage(..., raw_features=ShardedTable(...))` makes layer 1 gather through
    gs_agg_fwd_bf16_sharded (src/models.py:265,303-314 without ever assembling the table)."""

    def __init__(self, bases: Sequence[int], ld: int, rows_per_shard: int, num_nodes: int, dim: int, device,
                 local_shards: Optional[Sequence[Optional[torch.Tensor]]] = None, owner=None):
        self.lib = native.load()
        self.num_shards = len(bases)
        if not 1 <= self.num_shards <= 8:
            raise ValueError("1..8 shards (one NVSwitch box)")
        self.rows_per_shard, self.num_nodes, self.dim, self.ld = int(rows_per_shard), int(num_nodes), int(dim), int(ld)
        if self.ld % 8 or self.ld < self.dim:
            raise ValueError("bf16 rows must be ld % 8 == 0 elements apart (16-byte pieces)")
        if self.rows_per_shard * self.num_shards < self.num_nodes:
            raise ValueError("shards do not cover num_nodes")
        if any(int(b) == 0 or int(b) % 16 for b in bases):
            raise ValueError("shard base pointers must be non-null and 16-byte aligned")
        self.device = torch.device(device)
        self.bases = (ctypes.c_void_p * self.num_shards)(*[int(b) for b in bases])
        self.local_shards = list(local_shards) if local_shards is not None else [None] * self.num_shards
        self._owner = owner

    # what the reference reads from raw_features
    @property
    def shape(self):
        return (self.num_nodes, self.dim)

    def size(self, i: Optional[int] = None):
        return self.shape if i is None else self.shape[i]

    def __len__(self):
        return self.num_nodes

    @classmethod
    def from_full(cls, feats: torch.Tensor, num_shards: int) -> "ShardedTable":
        """Split a full [N, F] table into `num_shards` bf16 shards on the SAME device (tests and
        single-GPU use; the kernel cannot tell a local shard from a peer mapping)."""
        native.require_cuda(feats, "features")
        n, f = feats.shape
        rps = (n + num_shards - 1) // num_shards
        ld = (f + 7) & ~7
        shards = []
        for s in range(num_shards):
            t = torch.zeros((rps, ld), dtype=torch.bfloat16, device=feats.device)
            part = feats[s * rps:min(n, (s + 1) * rps)]
            t[:part.shape[0], :f] = part.to(torch.bfloat16)
            shards.append(t)
        return cls([t.data_ptr() for t in shards], ld, rps, n, f, feats.device, local_shards=shards)

    @classmethod
    def distributed(cls, local_rows: torch.Tensor, num_nodes: int, group=None) -> "ShardedTable":
        """Every rank passes its own shard (`local_rows`: [rows_per_shard, F], any float dtype,
        same shape on every rank).  The shard is copied into an IPC-exported region and the other
        ranks' regions are mapped; returns the table of `world` shards (peer shards are raw
        mapped addresses: only kernels of this library dereference them)."""
        import torch.distributed as dist
        native.require_cuda(local_rows, "local feature shard")
        rps, f = int(local_rows.shape[0]), int(local_rows.shape[1])
        ld = (f + 7) & ~7
        region = PeerRegion(rps * ld * 2, local_rows.device, group=group)
        mine = region.tensor(region.rank, 0, (rps, ld), torch.bfloat16)
        mine[:, :f] = local_rows.to(torch.bfloat16)
        torch.cuda.synchronize(local_rows.device)
        if region.world > 1:
            dist.barrier(group=group)                             # every shard is filled before anyone gathers
        local = [mine if r == region.rank else None for r in range(region.world)]
        return cls(region.ptrs, ld, rps, num_nodes, f, local_rows.device, local_shards=local, owner=region)

    def to_dense_fp32(self) -> torch.Tensor:
        """The full table as fp32 [N, F] (tests / oracle input); needs every shard as a local tensor."""
        if any(t is None for t in self.local_shards):
            raise RuntimeError("to_dense_fp32 needs all shards local (peer shards are raw mappings)")
        full = torch.cat([t[:, :self.dim].float() for t in self.local_shards], dim=0)
        return full[:self.num_nodes].contiguous()
