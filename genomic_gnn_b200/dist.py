"""Multi-GPU sharding of the graph build: one process per GPU, torch.distributed
(NCCL on the GPUs, gloo in the CPU tests) for the exchange steps.

    counting   shards by read chunk      -> all-reduce(sum) of the (k+1)-mer histogram,
                                            all-reduce(min) of the first-occurrence offsets,
                                            all-gather of the (rare) bridge records
    assembly   replicated                -> none
    KF         shards by 1024-row block  -> candidate-count sum, score-class histogram sum when the
                                            global top-n binds, all-gather(v) of the edge lists
  Nothing here touches a kernel: it is host logic over
tensors, which is why the CPU tests can cover it with gloo.
"""
from typing import List, Tuple

import torch
import torch.distributed as dist_


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``n_items`` (reads) over ranks."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def kf_block_ranges(n_blocks: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous row-block ranges with (nearly) equal upper-triangle work.
    Row block I scores blocks J >= I, i.e. n_blocks - I block pairs; contiguity
    keeps the concatenation of the shards in the reference's emission order."""
    prefix = [0]
    for i in range(n_blocks):
        prefix.append(prefix[-1] + (n_blocks - i))
    total = prefix[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        best = min(range(cuts[-1], n_blocks + 1), key=lambda i: (abs(prefix[i] - target), i))
        cuts.append(best)
    cuts.append(n_blocks)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def kf_block_range(n_blocks: int, world: int, rank: int) -> Tuple[int, int]:
    return kf_block_ranges(n_blocks, world)[rank]


class TorchComm:
    """The four exchange primitives the pipeline needs, over a torch.distributed group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist_.get_world_size(group)
        self.rank = dist_.get_rank(group)
        backend = dist_.get_backend(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    def sum_int(self, x: int) -> int:
        t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
        dist_.all_reduce(t, op=dist_.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def sum_tensor_(self, t: torch.Tensor) -> torch.Tensor:
        dist_.all_reduce(t, op=dist_.ReduceOp.SUM, group=self.group)
        return t

    def min_tensor_(self, t: torch.Tensor) -> torch.Tensor:
        dist_.all_reduce(t, op=dist_.ReduceOp.MIN, group=self.group)
        return t

    def gather_ints(self, x: int) -> List[int]:
        t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
        out = [torch.zeros_like(t) for _ in range(self.world)]
        dist_.all_gather(out, t, group=self.group)
        return [int(o.item()) for o in out]

    def gather_cat(self, t: torch.Tensor, dim: int = 0) -> torch.Tensor:
        """all-gather(v) along ``dim`` and concatenate in rank order."""
        sizes = self.gather_ints(t.shape[dim])
        mx = max(sizes)
        if mx == 0:
            return t
        tt = t.movedim(dim, 0).contiguous()
        pad = torch.zeros((mx,) + tuple(tt.shape[1:]), dtype=tt.dtype, device=tt.device)
        pad[: tt.shape[0]] = tt
        out = [torch.empty_like(pad) for _ in range(self.world)]
        dist_.all_gather(out, pad, group=self.group)
        cat = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
        return cat.movedim(0, dim).contiguous()


_I64_MAX = torch.iinfo(torch.int64).max


def reduce_counts(cr, comm: TorchComm):
    """C1: merge the per-rank K1 results so that every rank holds the global
    histogram, the global first occurrences and all bridge records."""
    from .core import CountResult

    hist = comm.sum_tensor_(cr.hist)  # uint32 counts in int32 storage: exact while totals < 2^31
    first = cr.first_pos
    first.masked_fill_(first < 0, _I64_MAX)  # GG_NO_POS is -1 as int64: lift it above every real offset
    first = comm.min_tensor_(first)
    n_br = comm.gather_ints(cr.n_bridges)
    bridges, n_bridges = cr.bridges, cr.n_bridges
    if sum(n_br) > 0:
        bridges = comm.gather_cat(cr.bridges[: cr.n_bridges].contiguous(), dim=0)
        n_bridges = int(bridges.shape[0])
        if n_bridges == 0:
            bridges = cr.bridges
    n_nonzero = int((hist != 0).sum().item())
    return CountResult(cr.k, hist, first, bridges, n_bridges, cr.status, n_nonzero)
