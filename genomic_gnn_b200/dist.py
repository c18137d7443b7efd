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
import ctypes as C
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


class PeerSpace:
    """NVLink peer memory for the latency-bound exchanges of a multi-GPU build (gg_peer_push / _signal / _wait).

    One symmetric allocation (torch.distributed._symmetric_memory: every rank's buffer is mapped into every rank's
    address space, plus a pad of 32-bit flags per rank), used as two halves that alternate from build to build;
    inside a build, regions are handed out by a bump allocator in the same order on every rank, so a region has the
    same offset everywhere.  ``all_gather(offset, seg_bytes)``: this rank has written segment ``rank`` of the region;
    on return (stream order) every segment is there.

    Why alternating halves are enough: a rank can only be pushing for exchange e of build s + 1 once every peer has
    signalled exchange e - 1 of build s + 1, i.e. once every peer has left build s behind on its own stream -- so
    nobody still reads a region of build s - 1, whose memory build s + 1 reuses."""

    CHANNELS = 16

    def __init__(self, comm, nbytes: int):
        import torch.distributed._symmetric_memory as symm

        from . import _lib

        self.comm = comm
        self.half = max(int(nbytes) // 2, 1 << 20) // 4096 * 4096
        group = comm.group if comm.group is not None else dist_.group.WORLD
        self.buf = symm.empty(2 * self.half, dtype=torch.uint8, device=comm.device)
        self.hdl = symm.rendezvous(self.buf, group)
        # buffer_ptrs are the bases of the symmetric allocations; this tensor starts hdl.offset bytes into them
        skew = int(getattr(self.hdl, "offset", 0) or 0)
        self.base = [int(p) + skew for p in self.hdl.buffer_ptrs]
        if self.base[comm.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric memory: own buffer pointer does not match the tensor")
        self.pads = [int(p) for p in self.hdl.signal_pad_ptrs]
        if int(self.hdl.signal_pad_size) < 4 * self.CHANNELS * comm.world:
            raise RuntimeError("signal pad too small")
        self.flag_peers = _lib.Peers(comm.world, (C.c_void_p * _lib.GG_MAX_PEERS)(
            *[None if r == comm.rank else self.pads[r] for r in range(comm.world)]))
        self.seq = [0] * self.CHANNELS
        self.parity, self.top, self.channel = 0, 0, 0
        self.status = torch.zeros(1, dtype=torch.int64, device=comm.device)
        self.used = 0

    def begin_build(self):
        self.parity ^= 1
        self.top, self.channel = 0, 0

    def alloc(self, nbytes: int):
        """Offset of a fresh region of this build, or None when the half is exhausted (callers fall back to NCCL;
        every rank takes the same decision because every rank asks for the same sizes)."""
        off = (self.top + 255) // 256 * 256
        if off + nbytes > self.half or self.channel >= self.CHANNELS:
            return None
        self.top = off + nbytes
        return self.parity * self.half + off

    def view(self, offset: int, nbytes: int, dtype) -> torch.Tensor:
        return self.buf[offset: offset + nbytes].view(dtype)

    def holds(self, t: torch.Tensor) -> bool:
        lo = self.buf.data_ptr()
        return lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= lo + 2 * self.half

    def all_gather(self, offset: int, seg_bytes: int):
        from . import _lib
        from .core import _stream, check

        lib = _lib.load()
        comm = self.comm
        if seg_bytes % 16:
            raise ValueError("segments must be multiples of 16 bytes")
        mine = offset + comm.rank * seg_bytes
        dst = _lib.Peers(comm.world, (C.c_void_p * _lib.GG_MAX_PEERS)(
            *[None if r == comm.rank else self.base[r] + mine for r in range(comm.world)]))
        ch = self.channel
        self.channel += 1
        self.seq[ch] += 1
        st = _stream()
        check(lib.gg_peer_push(C.c_void_p(self.base[comm.rank] + mine), seg_bytes, C.byref(dst), st), "gg_peer_push")
        check(lib.gg_peer_signal(C.byref(self.flag_peers), ch * comm.world + comm.rank, self.seq[ch], st), "gg_peer_signal")
        check(lib.gg_peer_wait(C.c_void_p(self.pads[comm.rank]), ch * comm.world, comm.world, comm.rank, self.seq[ch],
                               C.c_void_p(self.status.data_ptr()), st), "gg_peer_wait")
        self.used += 1

    def check(self):
        """Raise if a wait ever gave up on a peer (call at a point where the host synchronises anyway)."""
        if int(self.status.item()) != 0:
            raise RuntimeError("peer exchange: a rank never signalled (gg_peer_wait timed out)")


class TorchComm:
    """The four exchange primitives the pipeline needs, over a torch.distributed group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist_.get_world_size(group)
        self.rank = dist_.get_rank(group)
        backend = dist_.get_backend(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        self.peers = None

    def enable_peer_exchange(self, nbytes: int = 1 << 30) -> bool:
        """Route the small / latency-bound exchanges (count pieces, candidate counts, packed KF edges) through NVLink
        peer memory instead of NCCL.  Returns False -- and leaves NCCL in charge -- when symmetric memory cannot be
        set up on this box; the decision is taken together (all ranks or none)."""
        ok, err = 1, None
        try:
            self.peers = PeerSpace(self, nbytes)
        except Exception as exc:  # noqa: BLE001
            ok, err, self.peers = 0, exc, None
        flag = torch.tensor([ok], dtype=torch.int64, device=self.device)
        dist_.all_reduce(flag, op=dist_.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            self.peers = None
            self.peer_error = repr(err) if err is not None else "another rank failed"
            return False
        return True

    def begin_build(self):
        if self.peers is not None:
            self.peers.begin_build()

    def gather_buffer(self, n_elems_per_rank: int, dtype) -> torch.Tensor:
        """world * n_elems_per_rank elements for an in-place all-gather (all_gather_padded_): peer memory when it fits."""
        size = torch.empty(0, dtype=dtype).element_size()
        seg = n_elems_per_rank * size
        if self.peers is not None and seg % 16 == 0:
            off = self.peers.alloc(self.world * seg)
            if off is not None:
                return self.peers.view(off, self.world * seg, dtype)
        return torch.empty(self.world * n_elems_per_rank, dtype=dtype, device=self.device)

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
        from .core import read_small

        t = t.contiguous()
        n = int(t.numel())
        pad = (n + 1) // 2 * 2          # 16-byte segments
        out = self.gather_buffer(pad, torch.int64)
        if self.peers is not None and self.peers.holds(out):
            out[self.rank * pad: self.rank * pad + n].copy_(t)
            self.peers.all_gather(out.data_ptr() - self.peers.buf.data_ptr(), pad * 8)
        else:
            out = torch.empty(self.world * pad, dtype=torch.int64, device=t.device)
            mine = torch.zeros(pad, dtype=torch.int64, device=t.device)
            mine[:n] = t
            dist_.all_gather_into_tensor(out, mine, group=self.group)
        return read_small(out).view(self.world, pad)[:, :n].tolist()

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
        """out[rank * len(mine) : (rank + 1) * len(mine)] = every rank's ``mine`` (equal sizes).  When ``out`` is peer
        memory (gather_buffer) and ``mine`` already is its own segment of it, the exchange is a push over NVLink."""
        seg = mine.numel() * mine.element_size()
        if (self.peers is not None and self.peers.holds(out) and seg % 16 == 0
                and mine.data_ptr() == out.data_ptr() + self.rank * seg):
            self.peers.all_gather(out.data_ptr() - self.peers.buf.data_ptr(), seg)
            return out
        dist_.all_gather_into_tensor(out, mine, group=self.group)
        return out

    def all_gather_padded_(self, buf: torch.Tensor, stride: int):
        """In-place all-gather: ``buf`` holds world * stride elements and this rank has filled its own segment
        [rank * stride, (rank + 1) * stride)."""
        mine = buf[self.rank * stride : (self.rank + 1) * stride]
        return self.all_gather_into(buf, mine)

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
    everyone = getattr(piece, "_gg_everyone", None)   # core.count_piece: the piece is a segment of a gather buffer
    if everyone is None:
        everyone = torch.empty(world * length, dtype=torch.int64, device=piece.device)
    comm.all_gather_into(everyone, piece)
    hist = torch.empty(bins, dtype=torch.int32, device=piece.device)
    first = torch.empty(bins, dtype=torch.int64, device=piece.device)
    check(lib.gg_count_merge(ptr(everyone), world, length, cr.k, ptr(hist), ptr(first), _stream()), "gg_count_merge")
    n_tail = _lib.ST_WORDS + 1                    # status words + end of the rank's stream piece, right behind first_pos
    tail_at = bins // 2 + bins
    tails = read_small(everyone.view(world, length)[:, tail_at: tail_at + n_tail].contiguous()).tolist()
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
