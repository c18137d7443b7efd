"""Host delivery of a multi-GPU graph build: every rank copies ITS shard of the graph into its slice of one host
arena that all ranks of the box map (POSIX shared memory, page-locked with cudaHostRegister), so the consumer -- the
reference's CPU-side loaders in the launching process (kmer_ssgnn.py:162-230 hands CPU tensors to PyG) -- finds the
whole graph in host memory without any edge all-gather between the GPUs and with all PCIe links busy instead of one.

    arena = SharedHostArena(comm, nbytes)                 # once
    host = deliver_graph(built, arena, comm)              # per build: rank r writes its rows / its row-block shard
    host.kf_edge_index[i] ...                             # rank 0 (and every other rank) reads the assembled tensors

Layout of one delivery (identical on every rank, derived from the edge counts all ranks already share):
node_codes | DB edge_index | DB weight | x_0 .. | KF_0 edge_index | KF_0 weight | ...; a [2, E] edge_index is two rows
of E, and rank r owns the column range [sum(per_rank[:r]), sum(per_rank[:r + 1])) of both.
"""
import mmap
import os
import uuid
from typing import List

import numpy as np
import torch

from .core import BuiltGraph, HostGraph, _require_cuda


class SharedHostArena:
    """``nbytes`` of page-locked host memory shared by all ranks of ``comm`` (one box)."""

    def __init__(self, comm, nbytes: int, name: str = None):
        _require_cuda()
        self.comm, self.nbytes = comm, int(nbytes)
        names = [name or f"gg_b200_{os.getpid()}_{uuid.uuid4().hex[:8]}"]
        if comm is not None and comm.world > 1:
            import torch.distributed as dist

            dist.broadcast_object_list(names, src=0, group=comm.group)
        self.path = os.path.join("/dev/shm", names[0])
        rank = 0 if comm is None else comm.rank
        if rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.nbytes)
        self._barrier()
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self._map = mmap.mmap(fd, self.nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        os.close(fd)
        self.bytes = torch.from_numpy(np.frombuffer(self._map, dtype=np.uint8))
        rc = torch.cuda.cudart().cudaHostRegister(self.bytes.data_ptr(), self.nbytes, 0)
        if int(rc) != 0:
            raise RuntimeError(f"cudaHostRegister failed with {int(rc)}")
        self._barrier()
        if rank == 0:
            os.unlink(self.path)     # the mappings keep the memory alive; nothing is left behind in /dev/shm

    def _barrier(self):
        if self.comm is not None and self.comm.world > 1:
            import torch.distributed as dist

            dist.barrier(group=self.comm.group)

    def close(self):
        if self._map is not None:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaHostUnregister(self.bytes.data_ptr())
            self.bytes = None
            try:
                self._map.close()
            except BufferError:      # a view handed out earlier is still alive: the mapping goes with the process
                pass
            self._map = None


class _Layout:
    def __init__(self, arena: SharedHostArena):
        self.arena, self.off = arena, 0

    def take(self, shape, dtype) -> torch.Tensor:
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        self.off = (self.off + 255) // 256 * 256
        if self.off + n > self.arena.nbytes:
            raise RuntimeError(f"shared host arena too small: need {self.off + n} bytes, have {self.arena.nbytes}")
        t = self.arena.bytes[self.off: self.off + n].view(dtype).view(*shape)
        self.off += n
        return t


def arena_bytes_for(n_nodes: int, k: int, small_ks, edge_budgets: List[int]) -> int:
    """Upper bound of one delivery: DB edges <= 4 N, features N x 4^s float32, KF sets by their edge budgets."""
    total = 8 * n_nodes + 20 * 4 * n_nodes
    total += sum(4 * n_nodes * 4**int(s) for s in small_ks)
    total += sum(20 * int(e) for e in edge_budgets)
    return total + (1 << 20)   # every tensor starts on a 256-byte boundary: a megabyte of slack covers any layout


def deliver_graph(built: BuiltGraph, arena: SharedHostArena, comm) -> HostGraph:
    """Device -> shared host, sharded by rank.  ``built.kf`` must be sharded (build_graph(kf_gather=False)) or
    single-rank.  Replicated tensors are split by rows: rank r copies rows [r N / W, (r + 1) N / W) of every feature
    matrix; rank 0 copies the (small) De Bruijn graph.  Returns views of the arena; they are complete on every rank once
    the call returns (it ends with a barrier)."""
    world = 1 if comm is None else comm.world
    rank = 0 if comm is None else comm.rank
    lay = _Layout(arena)
    n, e_db = built.db.num_nodes, built.db.num_edges
    d2h = 0

    def put(dst, src):
        nonlocal d2h
        if src.numel():
            dst.copy_(src, non_blocking=True)
            d2h += src.numel() * src.element_size()

    h_codes = lay.take((n,), torch.int64)
    h_ei = lay.take((2, e_db), torch.int64)
    h_w = lay.take((e_db,), torch.float32)
    if rank == 0:
        put(h_codes, built.db.node_codes)
        put(h_ei, built.db.edge_index)
        put(h_w, built.db.weight)
    r0, r1 = n * rank // world, n * (rank + 1) // world
    h_x = []
    for x in built.x:
        hx = lay.take(tuple(x.shape), torch.float32)
        put(hx[r0:r1], x[r0:r1])
        h_x.append(hx)
    kf_ei, kf_w = [], []
    for kf in built.kf:
        per_rank = kf.per_rank if kf.per_rank is not None else [int(kf.edge_index.shape[1])]
        if world > 1 and not kf.sharded:
            raise ValueError("deliver_graph wants sharded KF edge sets (build_graph(..., kf_gather=False))")
        total, lo = sum(per_rank), sum(per_rank[:rank])
        hei = lay.take((2, total), torch.int64)
        hw = lay.take((total,), torch.float32)
        mine = int(kf.edge_index.shape[1])
        put(hei[0, lo: lo + mine], kf.edge_index[0])
        put(hei[1, lo: lo + mine], kf.edge_index[1])
        put(hw[lo: lo + mine], kf.weight)
        kf_ei.append(hei)
        kf_w.append(hw)
    torch.cuda.current_stream().synchronize()
    arena._barrier()
    return HostGraph(built.k, n, h_codes, h_ei, h_w, h_x, kf_ei, kf_w, 0, d2h)
