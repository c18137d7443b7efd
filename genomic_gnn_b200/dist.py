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
        out = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist_.all_gather_into_tensor(out, t, group=self.group)
        return [int(v) for v in out.tolist()]  # one device-to-host read for all ranks

    def gather_device_ints(self, t: torch.Tensor) -> List[List[int]]:
        """all-gather of a small int64 vector that already lives on the device: no upload, one read for all ranks.
        Returns one list per rank."""
        t = t.contiguous()
        out = torch.empty(self.world * t.numel(), dtype=torch.int64, device=t.device)
        dist_.all_gather_into_tensor(out, t, group=self.group)
        from .core import read_small

        return read_small(out).view(self.world, t.numel()).tolist()

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


    def all_gather_into(self, out: torch.Tensor, mine: torch.Tensor):
        """out[rank * len(mine) : (rank + 1) * len(mine)] = every rank's ``mine`` (equal sizes)."""
        dist_.all_gather_into_tensor(out, mine, group=self.group)
        return out

    def all_gather_padded_(self, buf: torch.Tensor, stride: int):
        """In-place all-gather: ``buf`` holds world * stride elements and this rank has filled its own segment
        [rank * stride, (rank + 1) * stride)."""
        mine = buf[self.rank * stride : (self.rank + 1) * stride]
        dist_.all_gather_into_tensor(buf, mine, group=self.group)
        return buf

    def exchange_slices(self, tensors, sizes):
        """Zero-copy all-gather(v): every tensor in ``tensors`` is 1-D, already sized for ALL ranks, and holds this
        rank's data in its own slice [offset(rank), offset(rank) + sizes[rank]).  One batched group of point-to-point
        sends / receives fills the other slices in place.  Returns the request handles (wait on them before use)."""
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        me = self.rank
        ops = []
        for t in tensors:
            mine = t[offs[me] : offs[me + 1]]
            for r in range(self.world):
                if r == me:
                    continue
                peer = dist_.get_global_rank(self.group, r) if self.group is not None else r
                if sizes[me] > 0:
                    ops.append(dist_.P2POp(dist_.isend, mine, peer, group=self.group))
                if sizes[r] > 0:
                    ops.append(dist_.P2POp(dist_.irecv, t[offs[r] : offs[r + 1]], peer, group=self.group))
        return dist_.batch_isend_irecv(ops) if ops else []

    def gather_packed(self, tensors, sizes, dim_list):
        """One all-gather for several ragged tensors: ``tensors[i]`` is ragged along ``dim_list[i]`` with
        ``sizes[r]`` entries on rank r.  Each rank packs its pieces into one byte buffer (padded to the largest
        rank), a single all_gather_into_tensor moves everything, and the pieces are unpacked in rank order."""
        mx = max(sizes)
        if mx == 0:
            return list(tensors)
        rows = [t.movedim(d, 0).contiguous() for t, d in zip(tensors, dim_list)]
        per_item = [r[0].numel() * r.element_size() if r.shape[0] else
                    (r.element_size() * int(torch.tensor(r.shape[1:]).prod()) if r.dim() > 1 else r.element_size())
                    for r in rows]
        offs = [0]
        for b in per_item:
            offs.append(offs[-1] + (mx * b + 15) // 16 * 16)
        mine = torch.zeros(offs[-1], dtype=torch.uint8, device=rows[0].device)
        for r, b, o in zip(rows, per_item, offs):
            n = r.shape[0]
            if n:
                mine[o : o + n * b] = r.reshape(-1).view(torch.uint8)
        flat = torch.empty(self.world * offs[-1], dtype=torch.uint8, device=mine.device)
        dist_.all_gather_into_tensor(flat, mine, group=self.group)
        everyone = flat.view(self.world, offs[-1])
        out = []
        for r, b, o, d in zip(rows, per_item, offs, dim_list):
            parts = [everyone[k, o : o + sizes[k] * b].view(r.dtype).reshape((sizes[k],) + tuple(r.shape[1:]))
                     for k in range(self.world)]
            out.append(torch.cat(parts, dim=0).movedim(0, d).contiguous())
        return out


_I64_MAX = torch.iinfo(torch.int64).max


def _reduce_counts_gathered(cr, comm: TorchComm, piece: torch.Tensor):
    """C1 on the GPU: ONE all-gather of every rank's (hist, first_pos, status, pos_limit) piece, a merge kernel
    (gg_count_merge) and one device read of the status words -- instead of two all-reduces and a dozen small tensor
    ops (0.31 ms of a 2.3 ms step at 4 GPUs)."""
    from . import _lib
    from .core import CountResult, _stream, check, ptr, read_small

    lib = _lib.load()
    world, length = comm.world, int(piece.numel())
    bins = int(cr.hist.numel())
    everyone = torch.empty(world * length, dtype=torch.int64, device=piece.device)
    comm.all_gather_into(everyone, piece)
    hist = torch.empty(bins, dtype=torch.int32, device=piece.device)
    first = torch.empty(bins, dtype=torch.int64, device=piece.device)
    check(lib.gg_count_merge(ptr(everyone), world, length, cr.k, ptr(hist), ptr(first), _stream()), "gg_count_merge")
    n_tail = _lib.ST_WORDS + 1
    tails = read_small(everyone.view(world, length)[:, length - n_tail:].contiguous()).tolist()
    bad = sum(int(t[_lib.ST_BAD_BYTES]) for t in tails)
    if bad > 0:
        raise ValueError(f"{bad} input bytes outside the alphabet ACGTN")
    n_br = [int(t[_lib.ST_BRIDGES]) for t in tails]
    capacity = int(cr.bridges.shape[0])
    if max(n_br) > capacity:
        raise OverflowError(max(n_br))  # the caller re-counts with room for every bridge record
    bridges, n_bridges = cr.bridges, n_br[comm.rank]
    if sum(n_br) > 0:
        (bridges,) = comm.gather_packed([cr.bridges[:n_bridges]], n_br, [0])
        n_bridges = int(bridges.shape[0])
    pos_limit = max(int(t[-1]) for t in tails)
    return CountResult(cr.k, hist, first, bridges, n_bridges, cr.status, -1, pos_limit=pos_limit)


def reduce_counts(cr, comm: TorchComm, piece: torch.Tensor = None):
    """C1: merge the per-rank K1 results so that every rank holds the global
    histogram, the global first occurrences and all bridge records.  With ``piece`` (core.count_piece, the buffer
    the rank's K1 outputs live in) the exchange is a single all-gather + merge kernel; without it (host-logic tests
    on gloo) it is two all-reduces."""
    from .core import CountResult

    if piece is not None and piece.is_cuda:
        return _reduce_counts_gathered(cr, comm, piece)

    # One SUM all-reduce carries the histogram and a small tail: every rank's bridge count (one slot per rank) and
    # the total of illegal bytes.  The tail is filled on the device from K1's status words, so the per-rank count was
    # launched without a host round trip (count_kp1mers(sync=False)) and this is the only one.
    n_bins = cr.hist.numel()
    world = comm.world
    packed = torch.zeros(n_bins + 3 * world + 1, dtype=cr.hist.dtype, device=cr.hist.device)
    packed[:n_bins] = cr.hist      # uint32 counts in int32 storage: exact while totals < 2^31
    packed[n_bins + comm.rank] = cr.status[1].to(packed.dtype)   # GG_ST_BRIDGES
    packed[n_bins + world] = cr.status[0].to(packed.dtype)       # GG_ST_BAD_BYTES
    # end of this rank's piece of the global stream (bounds K2's sort keys), as two 30-bit halves in its own slots
    limit = int(getattr(cr, "pos_limit", 1 << 40))
    packed[n_bins + world + 1 + 2 * comm.rank] = limit & 0x3FFFFFFF
    packed[n_bins + world + 2 + 2 * comm.rank] = limit >> 30
    comm.sum_tensor_(packed)
    hist = packed[:n_bins]
    first = cr.first_pos
    first.masked_fill_(first < 0, _I64_MAX)  # GG_NO_POS is -1 as int64: lift it above every real offset
    first = comm.min_tensor_(first)
    from .core import read_small

    tail = read_small(packed[n_bins:]).tolist()
    if tail[world] > 0:
        raise ValueError(f"{tail[world]} input bytes outside the alphabet ACGTN")
    n_br = [int(x) for x in tail[:world]]
    capacity = int(cr.bridges.shape[0])
    if max(n_br) > capacity:
        raise OverflowError(max(n_br))  # the caller re-counts with room for every bridge record
    bridges, n_bridges = cr.bridges, n_br[comm.rank]
    if sum(n_br) > 0:
        (bridges,) = comm.gather_packed([cr.bridges[:n_bridges]], n_br, [0])
        n_bridges = int(bridges.shape[0])
    # occupied-bin count: not worth a device round trip here; -1 = unknown, build_graph checks the built graph instead
    pos_limit = max(int(tail[world + 1 + 2 * r]) | (int(tail[world + 2 + 2 * r]) << 30) for r in range(world))
    return CountResult(cr.k, hist, first, bridges, n_bridges, cr.status, -1, pos_limit=pos_limit)
