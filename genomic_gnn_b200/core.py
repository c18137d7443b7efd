"""Device-level pipeline over the C ABI: torch supplies device memory and
streams, every computation is a kernel of libgg_b200.so.

Stages: count -> first occurrence -> De Bruijn assembly
-> sub-k-mer counts / features -> KF edges (threshold, global top-n).
"""
import ctypes as C
import os
import threading
from contextlib import contextmanager
from dataclasses import dataclass, field
from fractions import Fraction
from functools import lru_cache
from math import isqrt
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import GGError, KfParams, check, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class StageTimer:
    """CUDA-event timing of the pipeline stages on the launching stream (bench.py)."""

    def __init__(self):
        self.ranges = []

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.ranges:
            out.setdefault(name, []).append(a.elapsed_time(b))
        return out


_TIMER: Optional[StageTimer] = None


def set_timer(timer: Optional[StageTimer]):
    global _TIMER
    _TIMER = timer


_STAGE_PREFIX = threading.local()


@contextmanager
def _stage_prefix(prefix):
    """Stages timed inside are recorded under ``prefix + name`` (the chunked re-emission of a binding top-n is kept
    apart from the first scoring pass)."""
    old = getattr(_STAGE_PREFIX, "value", "")
    _STAGE_PREFIX.value = prefix
    try:
        yield
    finally:
        _STAGE_PREFIX.value = old


@contextmanager
def _stage(name):
    if _TIMER is None:
        yield
        return
    name = getattr(_STAGE_PREFIX, "value", "") + name
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _TIMER.ranges.append((name, a, b))


class LocalComm:
    """Single-process stand-in for the collectives of dist.TorchComm."""
    world, rank = 1, 0

    def sum_int(self, x):
        return int(x)

    def sum_tensor_(self, t):
        return t

    def gather_ints(self, x):
        return [int(x)]

    def gather_device_ints(self, t):
        return [read_small(t).tolist()]

    def gather_cat(self, t, dim=0):
        return t

    def gather_packed(self, tensors, sizes, dim_list):
        return list(tensors)

    def exchange_slices(self, tensors, sizes):
        return []

    def all_gather_padded_(self, buf, stride):
        return buf


_MAILBOX_TLS = threading.local()   # one mailbox per thread: a pinned buffer is not something two threads can share


def read_small(t: torch.Tensor) -> torch.Tensor:
    """Small device tensor -> host tensor through a persistent pinned mailbox (an async copy + one stream sync; a
    plain .cpu() stages through pageable memory and costs ~10 us more per read, and a build does five of them).
    The result is a view of the mailbox: consume it before the next read of the same dtype."""
    if not t.is_cuda:
        return t
    n = int(t.numel())
    key = (t.device, t.dtype)
    boxes = _MAILBOX_TLS.__dict__.setdefault("boxes", {})
    box = boxes.get(key)
    if box is None or box.numel() < n:
        box = torch.empty(max(n, 4096), dtype=t.dtype, pin_memory=True)
        boxes[key] = box
    box[:n].copy_(t.reshape(-1), non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return box[:n].view(t.shape)


class PendingRead:
    """A small device tensor on its way to the host: the copy is queued on the current stream now, ``get()`` waits for
    THAT copy only (an event), not for whatever the caller queued behind it.  A build hands the GPU its next kernels
    first and collects the sizes they do not depend on afterwards, so the device never idles across the read.
    Slots come from a per-thread ring of pinned buffers (a slot is reused after kSlots further reads)."""
    kSlots, kSlotBytes = 16, 8192

    def __init__(self, t: torch.Tensor):
        ring = _MAILBOX_TLS.__dict__.setdefault("ring", None)
        if ring is None:
            ring = _MAILBOX_TLS.ring = [torch.empty(self.kSlots * self.kSlotBytes, dtype=torch.uint8, pin_memory=True), 0]
        nbytes = int(t.numel()) * t.element_size()
        if nbytes > self.kSlotBytes:
            raise ValueError("PendingRead is for small tensors")
        slot = ring[1] % self.kSlots
        ring[1] += 1
        self.box = ring[0][slot * self.kSlotBytes: slot * self.kSlotBytes + nbytes].view(t.dtype).view(t.shape)
        self.box.copy_(t, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def get(self) -> torch.Tensor:
        """Host view of the data (valid until kSlots further reads have been issued by this thread)."""
        if self.event is not None:
            self.event.synchronize()
            self.event = None
        return self.box


def _require_cuda():
    if not torch.cuda.is_available():
        raise GGError("genomic_gnn_b200 needs a CUDA device: the graph-construction path has no CPU fallback")


# --------------------------------------------------------------------------- input marshalling
@dataclass
class ReadStream:
    """ASCII reads resident on the device.  fixed_len > 0: rows of that many
    bytes back to back; fixed_len == 0: one '\\n' after every read."""
    data: torch.Tensor  # uint8, padded to gg_stream_padded_bytes
    n_bytes: int
    fixed_len: int
    n_reads: int
    pos_base: int = 0


def _padded(n_bytes):
    return int(_lib.load().gg_stream_padded_bytes(n_bytes))


def pack_reads_host(sequences):
    """Host-side marshalling: -> (uint8 ndarray view, fixed_len, n_reads)."""
    if isinstance(sequences, np.ndarray):
        if sequences.dtype != np.uint8 or sequences.ndim != 2:
            raise ValueError("ndarray input must be uint8 [n_reads, read_len] ASCII")
        arr = np.ascontiguousarray(sequences)
        return arr.reshape(-1), int(arr.shape[1]), int(arr.shape[0])
    if isinstance(sequences, (bytes, bytearray)):
        raise ValueError("pass a list of reads, a uint8 [R, L] array or a ReadStream")
    seqs = list(sequences)
    if not seqs:
        return np.zeros(0, np.uint8), 0, 0
    joined = "\n".join(seqs) + "\n"
    return np.frombuffer(joined.encode("ascii"), dtype=np.uint8), 0, len(seqs)


def upload_reads(sequences, device=None, pos_base=0) -> ReadStream:
    _require_cuda()
    if isinstance(sequences, ReadStream):
        return sequences
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if isinstance(sequences, torch.Tensor):
        if sequences.dtype != torch.uint8 or sequences.dim() != 2:
            raise ValueError("tensor input must be uint8 [n_reads, read_len] ASCII")
        n_reads, fixed_len = int(sequences.shape[0]), int(sequences.shape[1])
        n_bytes = n_reads * fixed_len
        buf = torch.empty(_padded(n_bytes), dtype=torch.uint8, device=device)
        buf[:n_bytes].copy_(sequences.reshape(-1), non_blocking=True)
        return ReadStream(buf, n_bytes, fixed_len, n_reads, pos_base)
    flat, fixed_len, n_reads = pack_reads_host(sequences)
    n_bytes = int(flat.shape[0])
    buf = torch.empty(_padded(n_bytes), dtype=torch.uint8, device=device)
    if n_bytes:
        staging = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
        staging.numpy()[:] = flat
        buf[:n_bytes].copy_(staging, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # staging buffer is released on return
    return ReadStream(buf, n_bytes, fixed_len, n_reads, pos_base)


# --------------------------------------------------------------------------- K1
@dataclass
class CountResult:
    k: int
    hist: torch.Tensor        # uint32 as int32 storage [4^(k+1)]
    first_pos: torch.Tensor   # uint64 as int64 storage [4^(k+1)]
    bridges: torch.Tensor     # int32 [cap, 4] (u, v, pos_lo, pos_hi) = gg_bridge_t
    n_bridges: int
    status: torch.Tensor      # int64 [8]
    n_nonzero: int = -1
    pos_limit: int = 1 << 40  # every first-occurrence / bridge position is below this (bounds K2's radix sorts)


def count_piece(k: int, dev, comm=None) -> torch.Tensor:
    """One contiguous int64 buffer holding a rank's K1 outputs -- [hist: bins int32][first_pos: bins int64]
    [status: 8 int64][pos_limit] -- so that a multi-GPU build exchanges them with ONE all-gather (dist.reduce_counts).
    With a ``comm`` that hands out gather buffers the piece is this rank's segment of one (peer memory when enabled)."""
    bins = int(_lib.load().gg_hist_bins(k))
    length = (bins // 2 + bins + _lib.ST_WORDS + 1 + 1) // 2 * 2     # 16-byte segments
    if comm is not None and hasattr(comm, "gather_buffer"):
        everyone = comm.gather_buffer(length, torch.int64)
        piece = everyone[comm.rank * length: (comm.rank + 1) * length]
        piece._gg_everyone = everyone
        return piece
    return torch.empty(length, dtype=torch.int64, device=dev)


def count_kp1mers(stream: ReadStream, k: int, bridge_capacity: int = 1 << 16, first_occurrence: bool = True,
                  sync: bool = True, impl: str = "auto", piece: Optional[torch.Tensor] = None) -> CountResult:
    """K1 (+ first-occurrence scan).  Reference: src/utils.py:15-31,52-84.

    impl: "part" = keys partitioned into <= 8 streams of 15-bit keys, each counted with shared-memory atomics (k <= 8;
    k <= 6 counts in shared memory directly), "smem" = packed stream + per-CTA shared-memory histograms over slices of
    the bin space (k <= 8; superseded by "part", kept as an independent implementation), "l2" = one RED.ADD per window
    into the L2-resident table, "auto" = "part" when the table is small enough, else "l2"."""
    _require_cuda()
    lib = _lib.load()
    if not (1 <= k <= _lib.GG_MAX_K):
        raise ValueError(f"k must be in [1, {_lib.GG_MAX_K}] for the dense (k+1)-mer histogram")
    if impl not in ("auto", "part", "smem", "l2"):
        raise ValueError(f"unknown counting implementation {impl!r}")
    dev = stream.data.device
    bins = int(lib.gg_hist_bins(k))
    if piece is None:
        hist = torch.empty(bins, dtype=torch.int32, device=dev)
        first = torch.empty(bins, dtype=torch.int64, device=dev)
        status = torch.empty(_lib.ST_WORDS, dtype=torch.int64, device=dev)
    else:  # views into the caller's exchange buffer (count_piece)
        hist = piece[: bins // 2].view(torch.int32)
        first = piece[bins // 2: bins // 2 + bins]
        status = piece[bins // 2 + bins: bins // 2 + bins + _lib.ST_WORDS]
        piece[bins // 2 + bins + _lib.ST_WORDS: bins // 2 + bins + _lib.ST_WORDS + 1].fill_(stream.pos_base + stream.n_bytes)
    if impl == "auto":
        impl = "part" if int(lib.gg_count_part_workspace_bytes(stream.n_bytes, k)) else "l2"
    ws_bytes = 0
    if impl == "part":
        ws_bytes = int(lib.gg_count_part_workspace_bytes(stream.n_bytes, k))
    elif impl == "smem":
        ws_bytes = int(lib.gg_count_smem_workspace_bytes(stream.n_bytes, k))
    if impl != "l2" and ws_bytes == 0:
        raise ValueError(f"k={k}: the (k+1)-mer table is too large for the shared-memory counting paths")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    while True:
        bridges = torch.empty((max(bridge_capacity, 1), 4), dtype=torch.int32, device=dev)
        st = _stream()
        check(lib.gg_count_reset(k, ptr(hist), ptr(first), ptr(status), st), "gg_count_reset")
        with _stage("k1_count"):
            if impl == "l2":
                check(lib.gg_count_kp1mers(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, stream.pos_base,
                                           ptr(hist), ptr(bridges), bridge_capacity, ptr(status), st),
                      "gg_count_kp1mers")
            else:
                fn = lib.gg_count_kp1mers_part if impl == "part" else lib.gg_count_kp1mers_smem
                check(fn(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, stream.pos_base, ptr(hist),
                         ptr(bridges), bridge_capacity, ptr(status), ptr(ws), ws_bytes, st), "gg_count_kp1mers_" + impl)
        if first_occurrence:
            with _stage("k1_first"):
                check(lib.gg_first_occurrence(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, stream.pos_base,
                                              ptr(hist), ptr(first), ptr(status), st), "gg_first_occurrence")
        if not sync:
            return CountResult(k, hist, first, bridges, -1, status, pos_limit=stream.pos_base + stream.n_bytes)
        st_host = read_small(status).tolist()
        if int(st_host[_lib.ST_BAD_BYTES]) > 0:
            raise ValueError(f"{int(st_host[_lib.ST_BAD_BYTES])} input bytes outside the alphabet ACGTN "
                             "(the reference fails on these at gnn_utils.py:471)")
        nb = int(st_host[_lib.ST_BRIDGES])
        if nb <= bridge_capacity:
            return CountResult(k, hist, first, bridges, nb, status, int(st_host[_lib.ST_NONZERO]),
                               pos_limit=stream.pos_base + stream.n_bytes)
        bridge_capacity = nb  # rare: many internal N runs -- run again with room for all of them


def read_layout(sequences):
    """Host-side (start offset, length) of every read in the device stream that upload_reads builds."""
    if isinstance(sequences, (np.ndarray, torch.Tensor)):
        n, length = int(sequences.shape[0]), int(sequences.shape[1])
        return np.arange(n, dtype=np.int64) * length, np.full(n, length, dtype=np.int64)
    seqs = list(sequences)
    lens = np.fromiter(map(len, seqs), dtype=np.int64, count=len(seqs))
    starts = np.zeros(len(seqs), dtype=np.int64)
    if len(seqs):
        starts[1:] = np.cumsum(lens[:-1] + 1)  # one separator byte after every read
    return starts, lens


def count_pairs_stride(sequences, k: int, stride: int) -> CountResult:
    """Row f4: weighted_directed_edges(..., stride > 1, inlcudeUNK=False) (src/utils.py:52-84 with the windows of
    utils.py:28 at 0, stride, 2*stride, ...).  Every pair of consecutive surviving k-mers becomes one record; the
    (k+1)-mer histogram stays empty and gg_db_build de-duplicates the records."""
    _require_cuda()
    lib = _lib.load()
    if not (1 <= k <= _lib.GG_MAX_K) or stride < 1:
        raise ValueError("bad k / stride")
    stream = upload_reads(sequences)
    dev = stream.data.device
    starts, lens = read_layout(sequences)
    n = int(starts.shape[0])
    n_win = np.where(lens >= k, (lens - k) // stride + 1, 0)
    capacity = int(np.maximum(n_win - 1, 0).sum())  # upper bound: every window clean
    bins = int(lib.gg_hist_bins(k))
    hist = torch.zeros(bins, dtype=torch.int32, device=dev)
    first = torch.full((bins,), -1, dtype=torch.int64, device=dev)
    status = torch.zeros(_lib.ST_WORDS, dtype=torch.int64, device=dev)
    recs = torch.empty((max(capacity, 1), 4), dtype=torch.int32, device=dev)
    starts_d, lens_d = torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev)
    with _stage("k1_stride_pairs"):
        check(lib.gg_count_pairs_stride(ptr(stream.data), ptr(starts_d), ptr(lens_d), n, k, int(stride),
                                        stream.pos_base, ptr(recs), capacity, ptr(status), _stream()),
              "gg_count_pairs_stride")
    st = status.cpu()
    if int(st[_lib.ST_BAD_BYTES]) > 0:
        raise ValueError(f"{int(st[_lib.ST_BAD_BYTES])} input bytes outside the alphabet ACGTN")
    return CountResult(k, hist, first, recs, int(st[_lib.ST_BRIDGES]), status, 0,
                       pos_limit=stream.pos_base + stream.n_bytes)


# --------------------------------------------------------------------------- K2
@dataclass
class DBGraph:
    k: int
    num_nodes: int
    num_edges: int
    node_codes: torch.Tensor    # int64 [N] k-mer code of node i
    code_to_node: torch.Tensor  # int32 [4^k] (-1 = absent)
    edge_index: torch.Tensor    # int64 [2, E]
    weight: torch.Tensor        # float32 [E]
    weight64: torch.Tensor      # float64 [E]


class _NeedBridges(Exception):
    """build_db ran on a count whose status had not been read and found bridge records (args[0] of them)."""


def _db_launch(cr: "CountResult", normalise, create_all_kmers, edge_weight_threshold):
    """Queue gg_db_build; every output is sized by its upper bound (4^k nodes, 4^(k+1) + bridges edges) and the actual
    (N, E) stay on the device in ``ne``."""
    lib = _lib.load()
    dev = cr.hist.device
    k = cr.k
    bins = cr.hist.numel()
    cap = bins + cr.n_bridges
    ws_bytes = int(lib.gg_db_workspace_bytes(k, cr.n_bridges))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    node_codes = torch.empty(4**k, dtype=torch.int64, device=dev)
    code_to_node = torch.empty(4**k, dtype=torch.int32, device=dev)
    edges = torch.empty((2, cap), dtype=torch.int64, device=dev)
    w64 = torch.empty(cap, dtype=torch.float64, device=dev)
    w32 = torch.empty(cap, dtype=torch.float32, device=dev)
    ne = torch.empty(2, dtype=torch.int64, device=dev)
    thr = -1.0 if edge_weight_threshold is None else float(edge_weight_threshold)
    if edge_weight_threshold is not None and thr < 0:
        thr = 0.0 if thr == 0 else -1.0  # a negative threshold keeps every edge
    with _stage("k2_db"):
        check(lib.gg_db_build(ptr(cr.hist), ptr(cr.first_pos), ptr(cr.bridges), cr.n_bridges, k, int(bool(normalise)),
                              int(bool(create_all_kmers)), thr, ptr(node_codes), ptr(code_to_node), ptr(edges[0]),
                              ptr(edges[1]), ptr(w64), ptr(w32), cap, ptr(ne), ptr(ws), ws_bytes,
                              max(1, min(40, int(cr.pos_limit - 1).bit_length())), _stream()), "gg_db_build")
    return node_codes, code_to_node, edges, w64, w32, ne


def _db_check_speculation(cr: "CountResult", st_host) -> bool:
    """K1's status words, read after a De Bruijn assembly that assumed a stream without bridge records.  Raises on
    illegal bytes; True = bridge records exist and ``cr.n_bridges`` now says how many (assemble again)."""
    if int(st_host[_lib.ST_BAD_BYTES]) > 0:
        raise ValueError(f"{int(st_host[_lib.ST_BAD_BYTES])} input bytes outside the alphabet ACGTN "
                         "(the reference fails on these at gnn_utils.py:471)")
    cr.n_nonzero = int(st_host[_lib.ST_NONZERO])
    nb = int(st_host[_lib.ST_BRIDGES])
    if nb > 0:
        if nb > int(cr.bridges.shape[0]):
            cr.n_bridges = -1
            raise _NeedBridges(nb)            # the caller counts again with room for every bridge record
        cr.n_bridges = nb
        return True
    return False


def build_db(cr: CountResult, normalise=True, create_all_kmers=False, edge_weight_threshold=None,
             keep_top_k=None) -> DBGraph:
    """K2.  Reference: src/utils.py:87-165 + from_networkx.  keep_top_k = the quantile pruning of utils.py:151-160.

    A count whose status words are still on the device (``cr.n_bridges < 0``: count_kp1mers(sync=False)) is assembled
    on the assumption that it holds no bridge records -- reads without an internal N run have none -- and K1's status
    comes back in the SAME device read as K2's node / edge counts: one host round trip for both stages.  Illegal bytes
    raise here; bridge records make the call fix ``cr.n_bridges`` and assemble again (capacity permitting)."""
    lib = _lib.load()
    dev = cr.hist.device
    k = cr.k
    speculative = cr.n_bridges < 0
    if speculative:
        cr.n_bridges = 0
    node_codes, code_to_node, edges, w64, w32, ne = _db_launch(cr, normalise, create_all_kmers, edge_weight_threshold)
    if speculative:
        both = read_small(torch.cat([ne, cr.status])).tolist()
        n, e = int(both[0]), int(both[1])
        if _db_check_speculation(cr, both[2:]):
            return build_db(cr, normalise, create_all_kmers, edge_weight_threshold, keep_top_k)
    else:
        n, e = (int(x) for x in read_small(ne).tolist())
    ei, w32, w64 = edges[:, :e].contiguous(), w32[:e], w64[:e]
    if keep_top_k is not None and e > 0:
        pw = int(lib.gg_db_prune_workspace_bytes(e))
        pws = torch.empty(pw, dtype=torch.uint8, device=dev)
        o_ei = torch.empty((2, e), dtype=torch.int64, device=dev)
        o64 = torch.empty(e, dtype=torch.float64, device=dev)
        o32 = torch.empty(e, dtype=torch.float32, device=dev)
        cnt = torch.empty(1, dtype=torch.int64, device=dev)
        check(lib.gg_db_prune_quantile(ptr(ei[0]), ptr(ei[1]), ptr(w64), e, 1 - keep_top_k, ptr(o_ei[0]), ptr(o_ei[1]),
                                       ptr(o64), ptr(o32), ptr(cnt), ptr(pws), pw, _stream()), "gg_db_prune_quantile")
        e = int(cnt.item())
        ei, w32, w64 = o_ei[:, :e].contiguous(), o32[:e], o64[:e]
    return DBGraph(k, n, e, node_codes[:n], code_to_node, ei, w32, w64)


# --------------------------------------------------------------------------- K3
@dataclass
class KFCounts:
    k: int
    s: int
    counts: torch.Tensor       # uint8 [N, 4^s]
    sqnorm: torch.Tensor       # int32 [N]
    col_keep: torch.Tensor     # int32 [D'] columns that occur at least once
    @property
    def m(self):
        return self.k - self.s + 1


_ARANGE = {}


def _col_keep_from_host(present_host: np.ndarray, dev) -> torch.Tensor:
    """int32 indices of the columns that occur (gnn_utils.py:474-481 drops the others); the usual all-present case
    reuses a cached identity."""
    d = int(present_host.shape[0])
    if present_host.all():
        key = (d, str(dev))
        if key not in _ARANGE:
            _ARANGE[key] = torch.arange(d, dtype=torch.int32, device=dev)
        return _ARANGE[key]
    return torch.from_numpy(np.nonzero(present_host)[0].astype(np.int32)).to(dev)


def kf_counts_many(node_codes: torch.Tensor, k: int, small_ks) -> List[KFCounts]:
    """K3 for several small_k with ONE device read for all their column-presence flags."""
    lib = _lib.load()
    dev = node_codes.device
    n = int(node_codes.numel())
    ds = [4**int(s) for s in small_ks]
    present = torch.empty(sum(ds), dtype=torch.int32, device=dev)
    out, lo = [], 0
    for s, d in zip(small_ks, ds):
        counts = torch.empty((n, d), dtype=torch.uint8, device=dev)
        sqnorm = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        with _stage(f"k3_counts_d{d}"):
            check(lib.gg_kf_counts(ptr(node_codes), n, k, int(s), ptr(counts), ptr(sqnorm),
                                   C.c_void_p(present.data_ptr() + 4 * lo), _stream()), "gg_kf_counts")
        out.append(KFCounts(k, int(s), counts, sqnorm[:n], None))
        lo += d
    host = read_small(present).numpy().copy()
    lo = 0
    for kc, d in zip(out, ds):
        kc.col_keep = _col_keep_from_host(host[lo:lo + d], dev)
        lo += d
    return out


# counts matrices of the one-read flow are sized for 4^k rows before N is known: keep that for tables up to this size
_FUSED_COUNTS_BYTES = 2 << 30


def count_db_counts(stream: "ReadStream", k: int, small_ks, create_all_kmers=False):
    """K1 + K1b + K2 + K3 counts of a single-GPU build with ONE device read: the De Bruijn assembly assumes a stream
    without bridge records (build_db), the sub-k-mer counts are made for 4^k candidate rows with N read on the device
    (gg_kf_counts_upto), and N, E, K1's status words and every column-presence flag come back together.
    Returns (DBGraph, [KFCounts], CountResult).  Streams with bridge records, and tables too large to over-allocate,
    take the step-by-step path: the first two are None and the count comes back for build_db (its ``n_bridges`` says how
    many bridge records it holds; -1: more than it had room for, count again)."""
    lib = _lib.load()
    ds = [4**int(s) for s in small_ks]
    cr = count_kp1mers(stream, k, sync=False)
    if 4**k * sum(ds) > _FUSED_COUNTS_BYTES:
        return None, None, cr
    cr.n_bridges = 0
    node_codes, code_to_node, edges, w64, w32, ne = _db_launch(cr, True, create_all_kmers, None)
    dev = node_codes.device
    n_max = 4**k
    present = torch.empty(sum(ds), dtype=torch.int32, device=dev)
    raw, lo = [], 0
    for s_, d in zip(small_ks, ds):
        counts = torch.empty((n_max, d), dtype=torch.uint8, device=dev)
        sqnorm = torch.empty(n_max, dtype=torch.int32, device=dev)
        with _stage(f"k3_counts_d{d}"):
            check(lib.gg_kf_counts_upto(ptr(node_codes), n_max, ptr(ne), k, int(s_), ptr(counts), ptr(sqnorm),
                                        C.c_void_p(present.data_ptr() + 4 * lo), _stream()), "gg_kf_counts_upto")
        raw.append((int(s_), counts, sqnorm))
        lo += d
    both = read_small(torch.cat([ne, cr.status, present.view(torch.int64)]))
    head = both[: 2 + _lib.ST_WORDS].tolist()
    n, e = int(head[0]), int(head[1])
    try:
        if _db_check_speculation(cr, head[2:]):
            return None, None, cr          # bridge records: assemble with them, step by step
    except _NeedBridges:
        return None, None, cr
    host = both[2 + _lib.ST_WORDS:].numpy().view(np.int32).copy()
    db = DBGraph(k, n, e, node_codes[:n], code_to_node, edges[:, :e].contiguous(), w32[:e], w64[:e])
    kcs, lo = [], 0
    for (s_, counts, sqnorm), d in zip(raw, ds):
        kcs.append(KFCounts(k, s_, counts[:n], sqnorm[:n], _col_keep_from_host(host[lo:lo + d], dev)))
        lo += d
    return db, kcs, cr


def kf_counts(node_codes: torch.Tensor, k: int, s: int) -> KFCounts:
    """K3.  Reference: gnn_utils.py:454-481."""
    return kf_counts_many(node_codes, k, [int(s)])[0]


def value_lut(m: int) -> np.ndarray:
    """float64 feature value of count c: scipy scales the sparse count matrix by the
    reciprocal (gnn_utils.py:484-485), i.e. c * (1.0 / m), not c / m."""
    return np.arange(256, dtype=np.float64) * (1.0 / m)


def kf_features_f32(kc: KFCounts) -> torch.Tensor:
    """float32 [N, D'] node features as the GNN receives them
    (torch.tensor(labels, dtype=torch.float), gnn_tasks/utils.py:38)."""
    lib = _lib.load()
    dev = kc.counts.device
    n, d = kc.counts.shape
    dk = int(kc.col_keep.numel())
    lut = _device_value_lut(kc.m, dev)
    x = torch.empty((n, dk), dtype=torch.float32, device=dev)
    with _stage(f"k3_features_d{d}"):
        check(lib.gg_kf_features_f32(ptr(kc.counts), n, d, ptr(kc.col_keep), dk, ptr(lut), ptr(x), _stream()),
              "gg_kf_features_f32")
    return x


# --------------------------------------------------------------------------- K4-K6
@dataclass(frozen=True)
class ScoreTest:
    """The exact candidate test on integers: a pair with dot product ``dot`` and norm product P = sq_s * sq_t passes
    iff dot^2 * den > num * P (``inclusive``: >=).  A float64 threshold t is the rational num / den = t^2; the
    cut-off class (dot_c, P_c) of a binding top-n is num / den = dot_c^2 / P_c."""
    num: int
    den: int
    inclusive: bool = False

    @staticmethod
    def from_threshold(threshold: float) -> "ScoreTest":
        if threshold < 0:
            raise ValueError("negative KF threshold is not supported (the reference would then emit zero-weight "
                             "lower-triangle pairs)")
        fr = Fraction(threshold)
        return ScoreTest(fr.numerator**2, fr.denominator**2, False)


@lru_cache(maxsize=64)
def rational_tables(test: ScoreTest, sq_max: int):
    """Host tables of the scoring kernels (tiny).
    dmin[P]: smallest integer dot that passes ``test`` against the norm product P (65535: none can).
    fast_scale[sq]: conservative pre-filter of the wide-row kernels: a candidate always satisfies
    (dot^2 << shift) > fast_scale[sq_s] * sq_t."""
    a2, b2 = test.num, test.den
    p_max = sq_max * sq_max
    dmin = np.empty(p_max + 1, dtype=np.uint16)
    for p in range(p_max + 1):
        if test.inclusive:   # dot^2 >= ceil(a2 p / b2)
            need = -((-a2 * p) // b2)
            d = 0 if need <= 0 else isqrt(need - 1) + 1
        else:                # dot^2 > floor(a2 p / b2)
            d = isqrt((a2 * p) // b2) + 1
        dmin[p] = min(d, 65535)
    dmin[0] = 65535
    shift = 31 - max(p_max, 1).bit_length()
    shift = max(0, min(shift, 24))
    if test.inclusive:       # strictly below a2 sq 2^shift / b2, so that equality still passes the `>` of the kernels
        scale = np.array([max(((a2 * sq << shift) - 1) // b2, 0) if sq else 0 for sq in range(sq_max + 1)], dtype=np.uint64)
    else:
        scale = np.array([(a2 * sq << shift) // b2 for sq in range(sq_max + 1)], dtype=np.uint64)
    assert int(scale.max()) * sq_max < 2**32 and (p_max << shift) < 2**32
    return dmin, scale.astype(np.uint32), shift, p_max


def threshold_tables(threshold: float, sq_max: int):
    """Exact-rational tables of the strict test cos > threshold (gnn_utils.py:294)."""
    return rational_tables(ScoreTest.from_threshold(float(threshold)), int(sq_max))


_DEVICE_TABLES = {}


def _device_tables(test: ScoreTest, sq_max: int, dev):
    """Device copies of the tables, cached per (test, sq_max, device): the upload is a pageable host-to-device copy
    that would otherwise synchronise every call."""
    key = (test, sq_max, str(dev))
    hit = _DEVICE_TABLES.get(key)
    if hit is None:
        dmin, scale, shift, p_max = rational_tables(test, sq_max)
        hit = (torch.from_numpy(dmin.view(np.int16)).to(dev), torch.from_numpy(scale.view(np.int32)).to(dev), shift, p_max)
        if len(_DEVICE_TABLES) > 256:
            _DEVICE_TABLES.clear()
        _DEVICE_TABLES[key] = hit
    return hit


def _as_test(threshold) -> ScoreTest:
    return threshold if isinstance(threshold, ScoreTest) else ScoreTest.from_threshold(float(threshold))


_DEVICE_LUTS = {}


def _device_value_lut(m: int, dev):
    key = (m, str(dev))
    if key not in _DEVICE_LUTS:
        _DEVICE_LUTS[key] = torch.from_numpy(value_lut(m).astype(np.float32)).to(dev)
    return _DEVICE_LUTS[key]


def top_n_budget(n_nodes: int, keep_top_k: Optional[float], n_candidates: int) -> int:
    """gnn_utils.py:311-317, including that top_n == 0 keeps everything ([-0:])."""
    if keep_top_k is None:
        return n_candidates
    top_n = min(int(n_nodes**2 * keep_top_k), n_candidates)
    return n_candidates if top_n == 0 else top_n


@dataclass
class KFEdges:
    edge_index: torch.Tensor   # int64 [2, E], s < t, reference emission order (this rank's shard when ``sharded``)
    weight64: Optional[torch.Tensor]  # float64 [E] (None after a multi-GPU build or a streamed top-n)
    weight: torch.Tensor       # float32 [E]
    n_candidates: int          # pairs above the threshold before the top-n cut (all ranks)
    cutoff: Optional[float] = None   # score of the cut-off class when top-n was binding
    tie_kept: int = 0
    tie_total: int = 0
    impl: str = "auto"
    pending: list = field(default_factory=list)  # in-flight multi-GPU transfers filling edge_index / weight
    sharded: bool = False      # edge_index / weight hold only this rank's row-block shard (gather=False)
    per_rank: Optional[list] = None   # edges per rank, in rank (= emission) order
    resident: bool = True      # False: the edges were produced chunk by chunk into a recycled buffer (sink="discard")

    @property
    def n_edges(self) -> int:
        return int(sum(self.per_rank)) if self.per_rank is not None else int(self.edge_index.shape[1])

    def wait(self):
        for req in self.pending:
            req.wait()
        self.pending = []
        return self


def _row_blocks(n):
    return (n + _lib.GG_KF_BLOCK - 1) // _lib.GG_KF_BLOCK


def kf_hist_bins(sq_max: int) -> int:
    """Entries of a score-class table: class (dot, P) lives at dot * (p_max + 1) + P, dot <= sq_max, P <= sq_max^2."""
    return (sq_max + 1) * (sq_max * sq_max + 1)


def kf_sparse_index(counts: torch.Tensor, sqnorm: torch.Tensor):
    """Row lists + posting lists of a count matrix with sparse rows (gg_knn_sparse_index), for gg_kf_emit_sparse.
    Returns (index tensor, its size, device status int64[2]: [0] != 0 = a row has more than 16 non-zero counts)."""
    lib = _lib.load()
    dev = counts.device
    n, pitch = int(counts.shape[0]), int(counts.shape[1])
    ib = int(lib.gg_knn_sparse_index_bytes(n, pitch))
    index = torch.empty(ib, dtype=torch.uint8, device=dev)
    status = torch.empty(2, dtype=torch.int64, device=dev)
    key = ("nocls", str(dev))
    if key not in _DEVICE_LUTS:     # norm classes are a kNN matter: one empty table for every KF build
        _DEVICE_LUTS[key] = torch.full((256,), 0xFF, dtype=torch.uint8, device=dev)
    with _stage(f"k4_index_d{pitch}"):
        check(lib.gg_knn_sparse_index(ptr(counts), pitch, n, ptr(sqnorm), ptr(_DEVICE_LUTS[key]), 254, 1, ptr(index), ib,
                                      ptr(status), _stream()), "gg_knn_sparse_index")
    return index, ib, status


def kf_emit(counts: torch.Tensor, sqnorm: torch.Tensor, threshold, sq_max: int, iblock_begin=0,
            iblock_end=None, impl="auto", rec_capacity=None, sync=True, class_hist=None, hist_m=0, retry=True,
            sparse_index=None):
    """Score all pairs of the row-block shard and stage the candidates (``threshold``: float or ScoreTest).
    Returns (recs, n_recs, rowcnt, params, counts_scored): the last is the matrix the kernel read (zero-padded to a
    32-byte pitch for the tensor path on narrow rows), which kf_finalize needs to re-multiply narrow rows.
    With ``sync=False`` nothing is read back: n_recs is None and a sixth element, the device status words
    (gg_kf_emit: [0] candidates found, [2] pipeline error), is returned for the caller to read later.
    ``class_hist``: device uint64 table (kf_hist_bins entries) the CTA-pair tensor kernel adds the score classes of
    ALL candidates to; ``retry=False`` returns n_recs > capacity instead of scoring again with more room."""
    lib = _lib.load()
    dev = counts.device
    d_valid = int(counts.shape[1])
    if counts.shape[1] < 32 and impl in ("auto", "tcgen05"):
        # the tensor path wants >= 32 bytes of K per row (one kind::i8 MMA); zero columns change nothing
        counts = torch.nn.functional.pad(counts, (0, 32 - counts.shape[1]))
    n, d = int(counts.shape[0]), int(counts.shape[1])
    nb = _row_blocks(n)
    iblock_end = nb if iblock_end is None else iblock_end
    dmin_d, scale_d, shift, p_max = _device_tables(_as_test(threshold), int(sq_max), dev)
    sparse = impl == "sparse"
    if sparse and sparse_index is None:
        sparse_index = kf_sparse_index(counts, sqnorm)
    params = KfParams(n, d, iblock_begin, iblock_end, p_max, shift, _lib.KF_IMPL["auto" if sparse else impl], d_valid,
                      int(hist_m))
    rc_elems = int(lib.gg_kf_rowcnt_elems(n, iblock_begin, iblock_end))
    rowcnt = torch.empty(max(rc_elems, 1), dtype=torch.int64, device=dev)
    status = torch.empty(4, dtype=torch.int64, device=dev)
    if rec_capacity is None:
        rec_capacity = 1 << 22
    while True:
        recs = torch.empty((max(rec_capacity, 1), 4), dtype=torch.int32, device=dev)
        with _stage(f"k4_emit_d{d}"):
            if sparse:   # inverted-index join: only pairs that share a column can pass a threshold >= 0
                check(lib.gg_kf_emit_sparse(ptr(sparse_index[0]), sparse_index[1], d, n, iblock_begin, iblock_end,
                                            ptr(sqnorm), ptr(dmin_d), ptr(recs), rec_capacity, ptr(rowcnt), ptr(status),
                                            _stream()), "gg_kf_emit_sparse")
                status[3:4].copy_(sparse_index[2][0:1], non_blocking=True)   # [3] != 0: rows were not sparse after all
            else:
                check(lib.gg_kf_emit(C.byref(params), ptr(counts), ptr(sqnorm), ptr(dmin_d), ptr(scale_d), ptr(recs),
                                     rec_capacity, ptr(rowcnt), ptr(status), ptr(class_hist), _stream()), "gg_kf_emit")
        if not sync:
            return recs, None, rowcnt, params, counts, status
        st_host = read_small(status).tolist()
        if int(st_host[2]) != 0:
            raise GGError("KF tcgen05 kernel reported a pipeline time-out (mbarrier protocol error)")
        if sparse and int(st_host[3]) != 0:   # not k-mer count rows: the dense kernels take over
            return kf_emit(counts, sqnorm, threshold, sq_max, iblock_begin, iblock_end, "auto", rec_capacity, sync,
                           class_hist, hist_m, retry)
        found, slots = int(st_host[0]), max(int(st_host[0]), int(st_host[1]))
        if slots <= rec_capacity or not retry:
            out = _Emit((recs[:min(slots, rec_capacity)], found, rowcnt, params, counts))
            out.slots = slots
            return out
        del recs
        rec_capacity = slots + slots // 16 + (1 << 16)   # block reservations vary a little from run to run


class _Emit(tuple):
    """kf_emit's (recs, n_recs, rowcnt, params, counts_scored) + ``slots``: record slots the kernel used
    (>= n_recs: the CTA-pair kernel pads its slot blocks); more than the capacity means records were dropped."""
    slots = 0


def pair_kernel_serves(d: int, impl: str) -> bool:
    """True when gg_kf_emit runs the CTA-pair tensor kernel for rows of ``d`` bytes -- the kernel that can keep the
    score-class histogram while it scores."""
    return d % 128 == 0 and 128 <= d <= 1024 and impl in ("auto", "tcgen05x2")


class StagedRecords:
    """Candidates of a row-block shard as unordered records + per-(block pair, row, tile) counts (gg_kf_emit)."""

    def __init__(self, recs, n_recs, rowcnt, params, scored, status=None, redo=None):
        self.recs, self.n_recs, self.rowcnt, self.params, self.scored = recs, n_recs, rowcnt, params, scored
        self.d = params.d
        self.status4 = status      # device int64[4] while n_recs has not been read: [0] found, [2] pipeline error
        self.redo = redo           # re-scores the shard with room for a given number of records
        self.found = self.slots = None
        self.sparse = False        # scored by the inverted-index join (status word [3]: rows were not sparse)

    def after_counts(self, mine):
        """``mine`` = this rank's status words as read by kf_resolve."""
        if int(mine[2]) != 0:
            raise GGError("KF tcgen05 kernel reported a pipeline time-out (mbarrier protocol error)")
        self.found = int(mine[0])
        self.slots = max(self.found, int(mine[1]))
        if self.sparse and int(mine[3]) != 0:   # the sparse join met rows that are not sparse: score again, densely
            self.slots = int(self.recs.shape[0]) + 1
        self.status4 = None

    def complete(self):
        """Make sure every candidate has its record (the staging buffer may have been too small: score again)."""
        if self.n_recs is None:
            if self.slots > int(self.recs.shape[0]):
                self.recs, self.n_recs, self.rowcnt, self.params, self.scored = self.redo(self.slots + self.slots // 16 + (1 << 16))
            else:
                self.n_recs = self.found
                self.recs = self.recs[: self.slots]
        return self

    def place(self, ordered_ptr, capacity=None):
        """K6a: one packed 8-byte word per edge, in emission order, at ``ordered_ptr`` (room for n_recs words)."""
        if not self.n_recs:
            return
        lib = _lib.load()
        p = self.params
        pw = int(lib.gg_kf_place_workspace_bytes(p.n, p.iblock_begin, p.iblock_end))
        pws = torch.empty(pw, dtype=torch.uint8, device=self.recs.device)
        with _stage(f"k6_place_d{p.d}"):   # every record slot is scanned (padding records are skipped on the device)
            check(lib.gg_kf_place(C.byref(p), ptr(self.recs), int(self.recs.shape[0]), ptr(self.rowcnt), ptr(self.scored),
                                  ordered_ptr, ptr(pws), pw, _stream()), "gg_kf_place")


class StagedClasses:
    """Candidates of a row-block shard held as duplicate-row classes + the class-pair bit matrix (gg_kf_dedup_*)."""

    def __init__(self, counts, sqnorm, n, ib0, ib1, members, class_start, n_classes, adj, offsets, n_recs):
        self.counts, self.sqnorm, self.n, self.ib0, self.ib1 = counts, sqnorm, n, ib0, ib1
        self.members, self.class_start, self.n_classes, self.adj, self.offsets = members, class_start, n_classes, adj, offsets
        self.n_recs = n_recs                 # None until kf_resolve has read it
        self.d = int(counts.shape[1])
        self.status4 = torch.zeros(4, dtype=torch.int64, device=counts.device)
        self.status4[0:1].copy_(offsets[-1:], non_blocking=True)   # the shard's edge count, still on the device
        self.total_override = None

    def after_counts(self, mine):
        self.n_recs = int(mine[0])
        # [1] != 0: the scoring pass found more candidates in the whole matrix ([3]) than the top-n budget, and the
        # per-row counting pass was skipped on the device (its result would not be used)
        self.total_override = int(mine[3]) if int(mine[1]) != 0 else None
        self.status4 = None

    def complete(self):
        return self

    def score(self, test: ScoreTest, sq_max: int, extra=None):
        """(Re)build the class-pair bit matrix for ``test``; ``extra`` = _lib.KfDedupExtra by-products."""
        lib = _lib.load()
        dmin_d, _, _, p_max = _device_tables(test, int(sq_max), self.counts.device)
        with _stage(f"k4_score_d{self.d}"):
            check(lib.gg_kf_dedup_score(ptr(self.counts), self.d, self.d, ptr(self.sqnorm), ptr(self.members),
                                        ptr(self.class_start), self.n_classes, ptr(dmin_d), p_max, ptr(self.adj),
                                        None if extra is None else C.byref(extra), _stream()), "gg_kf_dedup_score")

    def count(self, ib0, ib1, skip_flag=None):
        """Exclusive prefix of the candidates per (block pair, row) of [ib0, ib1) under the current bit matrix; the
        last entry is their number.  ``skip_flag``: device int64, non-zero = leave everything zero."""
        lib = _lib.load()
        dev = self.counts.device
        elems = int(lib.gg_kf_rowcnt_elems(self.n, ib0, ib1))
        offsets = torch.empty(elems + 1, dtype=torch.int64, device=dev)
        cb = int(lib.gg_kf_dedup_count_workspace_bytes(self.n, ib0, ib1))
        cws = torch.empty(cb, dtype=torch.uint8, device=dev)
        with _stage(f"k4_count_d{self.d}"):
            check(lib.gg_kf_dedup_count(self.n, ib0, ib1, ptr(self.adj), ptr(self.members), ptr(self.class_start),
                                        self.n_classes, ptr(offsets), ptr(skip_flag), ptr(cws), cb, _stream()),
                  "gg_kf_dedup_count")
        return offsets

    def place_range(self, ib0, ib1, offsets, ordered_ptr, capacity):
        lib = _lib.load()
        ws = torch.empty(256, dtype=torch.uint8, device=self.counts.device)
        with _stage(f"k6_place_d{self.d}"):
            check(lib.gg_kf_dedup_place(self.n, ib0, ib1, ptr(self.adj), ptr(self.members), ptr(self.class_start),
                                        self.n_classes, ptr(offsets), ptr(self.counts), self.d, self.d, ordered_ptr,
                                        capacity, ptr(ws), 256, _stream()), "gg_kf_dedup_place")

    def place(self, ordered_ptr, capacity=None):
        if not self.n_recs:
            return
        self.place_range(self.ib0, self.ib1, self.offsets, ordered_ptr, self.n_recs if capacity is None else capacity)


def kf_stage_dedup(counts: torch.Tensor, sqnorm: torch.Tensor, threshold, sq_max: int, iblock_begin=0,
                   iblock_end=None, top_n_cap: Optional[int] = None) -> StagedClasses:
    """K4' for narrow rows (<= 16 columns, counts <= 15): group identical rows into classes, score the class pairs,
    count the shard's edges.  One small device read here (the number of classes sizes the bit matrix); the number of
    edges stays on the device for kf_resolve.  ``top_n_cap`` = int(N^2 p) of a global top-n: the scoring pass then
    also totals the candidates of the whole matrix (class pairs x member counts, no atomics per pair), and when they
    exceed the cap -- the top-n will cut, from the class table -- the per-row counting pass turns itself off on the
    device: at threshold 0 it would expand nearly every class against nearly every node for nothing."""
    return kf_stage_dedup_begin(counts, sqnorm, threshold, sq_max, iblock_begin, iblock_end, top_n_cap)()


def kf_stage_dedup_begin(counts: torch.Tensor, sqnorm: torch.Tensor, threshold, sq_max: int, iblock_begin=0,
                         iblock_end=None, top_n_cap: Optional[int] = None):
    """First half of kf_stage_dedup: the class kernels are queued and the class count is on its way to the host.
    Returns the second half (scoring + counting) as a callable; work queued between the two calls runs while the host
    waits for the class count."""
    lib = _lib.load()
    dev = counts.device
    n, d = int(counts.shape[0]), int(counts.shape[1])
    nb = _row_blocks(n)
    iblock_end = nb if iblock_end is None else iblock_end
    st = _stream()
    members = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    class_start = torch.empty(n + 1, dtype=torch.int32, device=dev)
    status = torch.empty(2, dtype=torch.int64, device=dev)
    wb = int(lib.gg_kf_dedup_classes_workspace_bytes(n))
    ws = torch.empty(wb, dtype=torch.uint8, device=dev)
    with _stage(f"k4_classes_d{d}"):
        check(lib.gg_kf_dedup_classes(ptr(counts), d, d, n, ptr(members), ptr(class_start), ptr(status), ptr(ws), wb, st),
              "gg_kf_dedup_classes")
    pending = PendingRead(status)

    def finish() -> StagedClasses:
        n_classes, n_over = (int(x) for x in pending.get().tolist())
        if n_over:
            raise GGError("duplicate-row KF path: a count above 15 does not fit the packed row key")
        adj = torch.empty(max(int(lib.gg_kf_dedup_adj_words(n_classes)), 1), dtype=torch.int32, device=dev)
        staged = StagedClasses.__new__(StagedClasses)
        staged.counts, staged.sqnorm, staged.n, staged.ib0, staged.ib1 = counts, sqnorm, n, iblock_begin, iblock_end
        staged.members, staged.class_start, staged.n_classes, staged.adj = members, class_start, n_classes, adj
        staged.d = d
        if top_n_cap:
            total = torch.zeros(2, dtype=torch.int64, device=dev)
            staged.score(_as_test(threshold), sq_max, _lib.KfDedupExtra(None, 0, 0, total.data_ptr(), 0, 0, 0, 0))
            binding = (total[0:1] > int(top_n_cap)).to(torch.int64)
            offsets = staged.count(iblock_begin, iblock_end, skip_flag=binding)
        else:
            staged.score(_as_test(threshold), sq_max)
            offsets = staged.count(iblock_begin, iblock_end)
        # the shard's edge count stays on the device: the caller reads it together with the other ranks' counts
        StagedClasses.__init__(staged, counts, sqnorm, n, iblock_begin, iblock_end, members, class_start, n_classes, adj,
                               offsets, None)
        if top_n_cap:
            staged.status4[1:2].copy_(binding, non_blocking=True)
            staged.status4[3:4].copy_(total[0:1], non_blocking=True)
        return staged

    return finish


def kf_finalize(staged, sqnorm, want_dot=False, out=None):
    """K6: packed placement + coalesced expansion.  ``out`` = (edge_index [2, >= n_recs] view, weight32 [>= n_recs]
    view) lets the caller have the edges written straight into its slice of a buffer sized for all ranks."""
    lib = _lib.load()
    dev = sqnorm.device
    n_recs = staged.n_recs
    if out is None:
        edges = torch.empty((2, max(n_recs, 1)), dtype=torch.int64, device=dev)
        w32 = torch.empty(max(n_recs, 1), dtype=torch.float32, device=dev)
    else:
        edges, w32 = out
    w64 = torch.empty(max(n_recs, 1), dtype=torch.float64, device=dev)
    dot = torch.empty(max(n_recs, 1), dtype=torch.int32, device=dev) if want_dot else None
    if n_recs:
        ordered = torch.empty(n_recs, dtype=torch.int64, device=dev)
        staged.place(ptr(ordered))
        first, size = (C.c_int64 * 1)(0), (C.c_int64 * 1)(n_recs)
        with _stage(f"k6_expand_d{staged.d}"):
            check(lib.gg_kf_expand(ptr(ordered), 1, first, size, ptr(sqnorm), ptr(edges[0]), ptr(edges[1]), ptr(w64),
                                   ptr(w32), ptr(dot), _stream()), "gg_kf_expand")
    return edges[:, :n_recs], w64[:n_recs], w32[:n_recs], (dot[:n_recs] if want_dot else None)


def choose_cutoff(class_hist: np.ndarray, p_max: int, budget: int):
    """Host side of K5: order the exact score classes (dot, P) by dot^2 / P as
    rationals, keep whole classes while they fit, and mark the class that
    straddles the budget.  Returns (action uint8 table, tie_budget, cutoff score,
    tie_total)."""
    action, tie_budget, cutoff, tie_total, _ = choose_cutoff_class(class_hist, p_max, budget)
    return action, tie_budget, cutoff, tie_total


def choose_cutoff_class(class_hist: np.ndarray, p_max: int, budget: int):
    """choose_cutoff + the score of the lowest class that contributes an edge, as the rational dot^2 / P: the cut-off
    class when one straddles the budget (tie_total > 0), else the last class kept whole."""
    keys = np.nonzero(class_hist)[0]
    dots, ps = keys // (p_max + 1), keys % (p_max + 1)
    cnt = class_hist[keys].astype(np.int64)
    vals = [Fraction(int(d) ** 2, int(p)) for d, p in zip(dots, ps)]
    order = sorted(range(len(keys)), key=lambda i: vals[i], reverse=True)
    action = np.zeros(class_hist.shape[0], dtype=np.uint8)
    taken, i = 0, 0
    tie_budget, cutoff, tie_total, cut = 0, None, 0, None
    while i < len(order) and taken < budget:
        j = i
        while j < len(order) and vals[order[j]] == vals[order[i]]:
            j += 1
        grp = [order[x] for x in range(i, j)]
        size = int(cnt[grp].sum())
        if taken + size <= budget:
            action[keys[grp]] = 1
            taken += size
            cut = vals[grp[0]]
        else:
            action[keys[grp]] = 2
            tie_budget, tie_total = budget - taken, size
            cutoff = float(dots[grp[0]]) / float(np.sqrt(float(ps[grp[0]])))
            cut = vals[grp[0]]
            taken = budget
        i = j
    return action, tie_budget, cutoff, tie_total, cut


def tie_ratio(tie_budget: int, tie_total: int) -> int:
    """r = ceil(B 2^64 / T) of the spread rule (gg_kf_select_packed): member g of the cut-off class, numbered in global
    emission order, is kept iff floor((g + 1) r / 2^64) > floor(g r / 2^64).  Exactly B of the T members pass."""
    if not (0 < tie_budget < tie_total):
        return 0
    return -((-tie_budget << 64) // tie_total)


def ties_admitted(g: int, ratio: int) -> int:
    """Members of the cut-off class with global number < g that the spread rule keeps."""
    return (g * ratio) >> 64


@dataclass
class KFStage:
    """One KF set between scoring and the read of its candidate count (kf_stage -> kf_resolve -> kf_finish)."""
    counts: torch.Tensor
    sqnorm: torch.Tensor
    threshold: float
    keep_top_k: Optional[float]
    sq_max: int
    impl: str
    staged: object = None          # StagedRecords / StagedClasses; None: no pair can be an edge
    per_rank: Optional[list] = None
    iblock_begin: int = 0
    iblock_end: int = 0
    row_sum: int = 0               # every row of counts sums to this (k - small_k + 1); 0: unknown
    class_hist: Optional[torch.Tensor] = None   # score classes of this rank's candidates, counted while scoring
    finish_stage: object = None    # kf_stage(defer=True): second half of a duplicate-row staging, not yet run


def kf_stage(counts: torch.Tensor, sqnorm: torch.Tensor, threshold: float = 0.0, keep_top_k: Optional[float] = None,
             sq_max: Optional[int] = None, impl: str = "auto", iblock_begin=0, iblock_end=None,
             row_sum: int = 0, defer: bool = False) -> KFStage:
    """K4: score this rank's row-block shard; the candidate count stays on the device (kf_resolve reads the counts of
    any number of staged sets, on every rank, with one device read).  ``defer``: a set that goes through duplicate-row
    classes only queues its class kernels; ``st.finish_stage()`` (set when deferred) waits for the class count and queues
    the rest -- whatever the caller queues in between runs while the host waits."""
    _require_cuda()
    n = int(counts.shape[0])
    if threshold < 0:
        raise ValueError("negative KF threshold is not supported")
    if sq_max is None:
        sq_max = int(sqnorm.max().item()) if n else 0
    iblock_end = _row_blocks(n) if iblock_end is None else iblock_end
    st = KFStage(counts, sqnorm, float(threshold), keep_top_k, int(sq_max), impl, iblock_begin=iblock_begin,
                 iblock_end=iblock_end, row_sum=int(row_sum))
    if n < 2 or threshold >= 1.0 or sq_max == 0:
        return st
    d_valid = int(counts.shape[1])
    # duplicate-row classes pay off once the pair count dwarfs their fixed cost (sort + two class sweeps, ~0.3 ms):
    # measured 0.96 vs 1.48 ms at N = 65,536 (k=8) but 0.32 vs 0.15 ms at N = 16,384 (k=7)
    use_dedup = impl == "dedup" or (impl == "auto" and d_valid <= 16 and n >= 32768)
    if use_dedup and not (d_valid <= 16 and d_valid % 4 == 0 and n <= _lib.KF_DEDUP_MAX_NODES and sq_max <= 225):
        if impl == "dedup":
            raise ValueError("the duplicate-row KF path needs <= 16 columns, counts <= 15 and <= 2^20 nodes")
        use_dedup = False
    if use_dedup:
        cap = int(n**2 * keep_top_k) if keep_top_k is not None else None
        second_half = kf_stage_dedup_begin(counts, sqnorm, threshold, sq_max, iblock_begin, iblock_end, top_n_cap=cap)
        if defer:
            def finish_stage():
                st.staged = second_half()
                st.finish_stage = None

            st.finish_stage = finish_stage
        else:
            st.staged = second_half()
    else:
        # staging capacity: the global top-n budget is the natural first guess (a binding threshold stays below it,
        # a binding top-n overflows, and is then cut from the score-class table without these records)
        n_pairs = n * (n - 1) // 2
        guess = int(n**2 * keep_top_k) if keep_top_k is not None else 1 << 22
        # ... capped at 2 GB of records: at k=10 the budget alone would be 1.1e10 records (176 GB)
        rec_capacity = max(1 << 16, min(n_pairs, max(guess, 1 << 20), 1 << 27))
        # impl="sparse": wide rows of k-mer counts are sparse (at most row_sum non-zero entries), and the inverted-index
        # join scores only the pairs that share a column -- 1.5 % of them at small_k = 5 -- up to 200,000 nodes (one
        # accumulator byte per node in shared memory).  On request only: at BASELINE config 4 it takes 1.28 + 0.12 ms
        # (one row per CTA iteration, latency-bound) against 1.15 ms for the CTA-pair tensor kernel, which stays the default
        lib = _lib.load()
        use_sparse = impl == "sparse"
        if use_sparse and not lib.gg_knn_sparse_eligible(d_valid, n, 1, 1):
            raise ValueError("the sparse KF path needs 64..1024 columns and at most 200,000 nodes")
        emit_impl = "sparse" if use_sparse else impl
        sparse_index = kf_sparse_index(counts, sqnorm) if use_sparse else None
        if keep_top_k is not None and not use_sparse and pair_kernel_serves(d_valid, impl):
            st.class_hist = torch.zeros(kf_hist_bins(int(sq_max)), dtype=torch.int64, device=counts.device)
        recs, _, rowcnt, params, scored, status = kf_emit(counts, sqnorm, threshold, sq_max, iblock_begin, iblock_end,
                                                          emit_impl, rec_capacity=rec_capacity, sync=False,
                                                          class_hist=st.class_hist, hist_m=row_sum,
                                                          sparse_index=sparse_index)

        def redo(room):
            # (a sparse staging whose rows turned out dense comes back here too: kf_emit then falls back by itself)
            return kf_emit(counts, sqnorm, threshold, sq_max, iblock_begin, iblock_end, emit_impl, rec_capacity=room,
                           sparse_index=sparse_index)

        st.staged = StagedRecords(recs, None, rowcnt, params, scored, status, redo)
        st.staged.sparse = use_sparse
        if use_sparse:
            st.impl = "auto"   # what a later re-emission of this set (binding top-n, chunked) may use
    return st


def kf_resolve(stages: List[KFStage], comm=None):
    """One exchange + one device read for the candidate counts (and kernel status words) of all staged KF sets."""
    comm = comm or LocalComm()
    live = [st for st in stages if st.staged is not None]
    if not live:
        return
    flat = live[0].staged.status4 if len(live) == 1 else torch.cat([st.staged.status4 for st in live])
    with _stage("c3_gather_counts"):
        table = comm.gather_device_ints(flat)      # [rank][4 * set + word]
    for i, st in enumerate(live):
        st.staged.after_counts(table[comm.rank][4 * i: 4 * i + 4])
        st.per_rank = [int(table[r][4 * i]) for r in range(comm.world)]


def kf_sets_overlapped(kcs, threshold: float, keep_top_k: Optional[float], impl: str = "auto", between=None) -> List["KFEdges"]:
    """Every KF set of a single-GPU build with no idle device across the three size reads it needs (class count of the
    duplicate-row sets, candidate counts of the record sets, edge counts of the duplicate-row sets): each read is a
    PendingRead queued BEHIND work that does not depend on it, so the host collects it while the device is busy.
    Order on the stream: scoring of the wide sets | class kernels of the narrow sets | placement of the wide sets |
    class-pair scoring + counting of the narrow sets | ``between()`` (the caller's independent kernels: float features)
    | placement of the narrow sets.  Same results as kf_stage + kf_resolve + kf_finish set by set."""
    stages = [None] * len(kcs)

    def stage(indices):
        """Queue the scoring of these sets; returns (sets staged as records, their pending candidate counts)."""
        for i in indices:
            kc = kcs[i]
            stages[i] = kf_stage(kc.counts, kc.sqnorm, threshold, keep_top_k, sq_max=kc.m * kc.m, impl=impl,
                                 row_sum=kc.m, defer=True)
        recs = [stages[i] for i in indices if stages[i].staged is not None]
        return recs, (PendingRead(torch.cat([st.staged.status4 for st in recs])) if recs else None)

    done = {}

    def finish(group, pending):
        if pending is None:
            return
        table = pending.get().tolist()
        for j, st in enumerate(group):
            st.staged.after_counts(table[4 * j: 4 * j + 4])
            st.per_rank = [int(table[4 * j])]
        for st in group:
            done[id(st)] = kf_finish(st)

    # wide sets first: their scoring kernels are the long ones, everything else queues up behind them
    wide, p_wide = stage([i for i in range(len(kcs)) if int(kcs[i].counts.shape[1]) > 16])
    small, p_small = stage([i for i in range(len(kcs)) if int(kcs[i].counts.shape[1]) <= 16])
    narrow = [st for st in stages if st.finish_stage is not None]
    finish(wide, p_wide)                   # placement + expansion of the wide sets queued
    for st in narrow:                      # class counts are in: class-pair scoring + counting queued
        st.finish_stage()
    p_narrow = PendingRead(torch.cat([st.staged.status4 for st in narrow])) if narrow else None
    finish(small, p_small)                 # narrow sets that went through the record path (few nodes)
    if between is not None:
        between()                          # the caller's kernels cover the read of the narrow sets' edge counts
    finish(narrow, p_narrow)
    return [done[id(st)] if id(st) in done else kf_finish(st) for st in stages]


def kf_edges(counts: torch.Tensor, sqnorm: torch.Tensor, threshold: float = 0.0, keep_top_k: Optional[float] = None,
             sq_max: Optional[int] = None, impl: str = "auto", iblock_begin=0, iblock_end=None, comm=None,
             gather: bool = True, row_sum: int = 0, chunk_records: Optional[int] = None, sink: str = "device") -> KFEdges:
    """K4-K6.  Reference: gnn_utils.py:240-322.

    With ``comm`` (dist.TorchComm) every rank scores its own row-block shard
    [iblock_begin, iblock_end); the only exchanges are an all-gather of the candidate counts,
    (when the global top-n binds) a sum of the score-class table, and -- unless ``gather=False`` leaves the result
    sharded -- the final all-gather of the per-shard edge lists, which concatenate in rank
    order to the single-GPU emission order."""
    st = kf_stage(counts, sqnorm, threshold, keep_top_k, sq_max, impl, iblock_begin, iblock_end, row_sum)
    kf_resolve([st], comm)
    return kf_finish(st, comm, gather, chunk_records=chunk_records, sink=sink)


def _block_pairs_rows(n, b0, b1):
    """Node pairs s < t whose smaller node lies in row blocks [b0, b1)."""
    lo, hi = min(b0 * _lib.GG_KF_BLOCK, n), min(b1 * _lib.GG_KF_BLOCK, n)
    rows = hi - lo
    return rows * (n - 1) - (lo + hi - 1) * rows // 2


class _SelectSink:
    """Destination of a top-n selection that arrives chunk by chunk in emission order: final edge arrays sized for the
    edges this rank keeps ("device"), or one recycled chunk-sized buffer when the caller only wants the work done and
    counted ("discard": the result would not fit the device, e.g. BASELINE config 5 on fewer than 4 GPUs)."""

    def __init__(self, dev, n_out, g_start, want_w64, discard, chunk_cap, out=None):
        self.discard = discard
        size = max(min(n_out, chunk_cap) if discard else n_out, 1)
        if out is not None:
            self.ei, self.w32 = out
        else:
            self.ei = torch.empty((2, size), dtype=torch.int64, device=dev)
            self.w32 = torch.empty(size, dtype=torch.float32, device=dev)
        self.w64 = torch.empty(size, dtype=torch.float64, device=dev) if want_w64 and not discard else None
        self.capacity = int(self.w32.shape[0])
        self.cursors = torch.tensor([int(g_start), 0], dtype=torch.int64, device=dev)
        self.kept_before = torch.zeros(1, dtype=torch.int64, device=dev)

    def select(self, ordered, n_words, sqnorm, p_max, action_d, ratio):
        lib = _lib.load()
        if n_words == 0:
            return
        if self.discard:   # recycle the buffer: remember what earlier chunks kept, start again at its front
            self.kept_before += self.cursors[1:2]
            self.cursors[1:2].zero_()
        wb = int(lib.gg_kf_select_packed_workspace_bytes(n_words))
        ws = torch.empty(wb, dtype=torch.uint8, device=ordered.device)
        with _stage("k5_select"):
            check(lib.gg_kf_select_packed(ptr(ordered), n_words, ptr(sqnorm), p_max, ptr(action_d), C.c_uint64(ratio),
                                          ptr(self.cursors), ptr(self.ei[0]), ptr(self.ei[1]), ptr(self.w64),
                                          ptr(self.w32), self.capacity, ptr(ws), wb, _stream()), "gg_kf_select_packed")

    def totals(self):
        """(members of the cut-off class seen up to here in global numbering, edges kept by this rank); ``last`` =
        edges sitting in the buffer now (all of them, or the last chunk's when the buffer is recycled)."""
        g, kept = read_small(self.cursors).tolist()
        self.last = min(int(kept), self.capacity)
        return int(g), int(kept) + (int(read_small(self.kept_before)[0]) if self.discard else 0)


def _plan_chunks(n, b0, b1, expect_words, cap):
    """Split row blocks [b0, b1) into consecutive ranges whose expected number of packed words stays below ``cap``
    (uniform density over the shard's node pairs; a range that still overflows is halved by the caller)."""
    pairs = max(_block_pairs_rows(n, b0, b1), 1)
    density = expect_words / pairs
    out, lo = [], b0
    while lo < b1:
        hi = lo + 1
        while hi < b1 and _block_pairs_rows(n, lo, hi + 1) * density * 1.25 <= cap:   # slot blocks are padded: ~12 % slack
            hi += 1
        out.append((lo, hi))
        lo = hi
    return out


def kf_finish(st: KFStage, comm=None, gather: bool = True, chunk_records: Optional[int] = None,
              sink: str = "device") -> KFEdges:
    """K5/K6 (+ the multi-GPU exchange) of a staged and resolved KF set.

    Threshold-bound sets (candidates <= int(N^2 p), or no top-n at all) are placed and expanded as they were scored.
    When the global top-n binds (gnn_utils.py:311-317) the cut-off comes from the table of exact score classes
    (dot, sq_s sq_t): counted inside the scoring kernels when they can (CTA-pair tensor kernel, duplicate-row classes)
    -- then the candidates are never materialised: the shard is re-scored in row-block chunks against the cut-off,
    >= for the cut-off class, and gg_kf_select_packed keeps what survives -- or from the placed candidates otherwise.
    Members of the cut-off class are kept by the spread rule (tie_ratio): the same members for every world size and
    chunking, an equal share on every rank."""
    lib = _lib.load()
    comm = comm or LocalComm()
    counts, sqnorm, threshold, keep_top_k, sq_max, impl = st.counts, st.sqnorm, st.threshold, st.keep_top_k, st.sq_max, st.impl
    dev = counts.device
    n = int(counts.shape[0])
    if sink not in ("device", "discard"):
        raise ValueError("sink must be 'device' or 'discard'")
    if st.staged is None:
        return KFEdges(torch.zeros((2, 0), dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.float64, device=dev),
                       torch.zeros(0, dtype=torch.float32, device=dev), 0, impl=impl, per_rank=[0] * comm.world,
                       sharded=comm.world > 1 and not gather)
    staged, per_rank = st.staged, st.per_rank
    n_total = sum(per_rank)
    if getattr(staged, "total_override", None) is not None:   # duplicate-row path: counted per class pair, no shard counts
        n_total = staged.total_override
    budget = top_n_budget(n, keep_top_k, n_total)
    need_cut = budget < n_total
    exchange = gather and comm.world > 1
    if need_cut:
        return _kf_finish_top_n(st, comm, gather, budget, n_total, chunk_records, sink)
    staged.complete()
    n_recs = staged.n_recs

    def shared_buffers(sizes):
        """Output tensors sized for every rank; this rank's slice starts at `lo`."""
        total, lo = sum(sizes), sum(sizes[: comm.rank])
        return (torch.empty((2, max(total, 1)), dtype=torch.int64, device=dev),
                torch.empty(max(total, 1), dtype=torch.float32, device=dev), total, lo)

    if exchange:
        # C2 on the packed form: K6a places this rank's edges as 8-byte words into its segment of an all-ranks
        # buffer, ONE in-place all-gather moves 8 instead of 20 bytes per edge, and K6b expands the gathered words
        # into edge_index / weight on every rank.  (float64 weights are not produced here: only the function-level
        # API wants them.)
        stride = (max(max(per_rank), 1) + 1) // 2 * 2        # 16-byte segments
        ordered = (comm.gather_buffer(stride, torch.int64) if hasattr(comm, "gather_buffer")
                   else torch.empty(comm.world * stride, dtype=torch.int64, device=dev))
        staged.place(C.c_void_p(ordered.data_ptr() + 8 * comm.rank * stride))
        with _stage("c2_gather_packed_edges"):
            comm.all_gather_padded_(ordered, stride)
        full_ei, full_w, e_total, lo = shared_buffers(per_rank)
        seg_first = (C.c_int64 * comm.world)(*[r * stride for r in range(comm.world)])
        seg_size = (C.c_int64 * comm.world)(*per_rank)
        with _stage(f"k6_expand_d{staged.d}"):
            check(lib.gg_kf_expand(ptr(ordered), comm.world, seg_first, seg_size, ptr(sqnorm), ptr(full_ei[0]),
                                   ptr(full_ei[1]), None, ptr(full_w), None, _stream()), "gg_kf_expand")
        return KFEdges(full_ei[:, :e_total], None, full_w[:e_total], n_total, None, 0, 0, impl, per_rank=list(per_rank))
    ei, w64, w32, _ = kf_finalize(staged, sqnorm)
    return KFEdges(ei.contiguous(), w64, w32, n_total, None, 0, 0, impl, per_rank=list(per_rank),
                   sharded=comm.world > 1)


def _kf_finish_top_n(st: KFStage, comm, gather, budget, n_total, chunk_records, sink) -> KFEdges:
    """The binding global top-n (see kf_finish)."""
    lib = _lib.load()
    counts, sqnorm, sq_max, impl, staged = st.counts, st.sqnorm, st.sq_max, st.impl, st.staged
    dev = counts.device
    n = int(counts.shape[0])
    b0, b1 = st.iblock_begin, st.iblock_end
    test = ScoreTest.from_threshold(st.threshold)
    p_max = sq_max * sq_max
    bins = kf_hist_bins(sq_max)
    dedup = isinstance(staged, StagedClasses)
    streamed = dedup or st.class_hist is not None
    chunk_cap = int(chunk_records) if chunk_records else 1 << 27
    ordered_all = None
    # ---- 1. the global table of score classes
    if dedup:
        # class pairs weighted by their member counts; every rank scores its slice of the class-pair triangle
        hist = torch.zeros(bins, dtype=torch.int64, device=dev)
        u = staged.n_classes
        cut = [int(round(u * (1.0 - (1.0 - r / comm.world) ** 0.5))) for r in range(comm.world)] + [u]
        extra = _lib.KfDedupExtra(hist.data_ptr(), cut[comm.rank], cut[comm.rank + 1], None, 0, 0, 0, 0)
        staged.score(test, sq_max, extra)
        local_hist = None
    elif st.class_hist is not None:
        hist = st.class_hist
        local_hist = read_small(hist).numpy().copy() if comm.world > 1 else None
    else:
        # candidates materialised (any other scoring kernel): place them and count the classes of the placed words
        staged.complete()
        ordered_all = torch.empty(max(staged.n_recs, 1), dtype=torch.int64, device=dev)
        staged.place(ptr(ordered_all))
        hist = torch.zeros(bins, dtype=torch.int64, device=dev)
        with _stage("k5_class_hist"):
            check(lib.gg_kf_class_hist_packed(ptr(ordered_all), staged.n_recs, ptr(sqnorm), p_max, ptr(hist), _stream()),
                  "gg_kf_class_hist_packed")
        local_hist = read_small(hist).numpy().copy() if comm.world > 1 else None
    if comm.world > 1:
        with _stage("c5_reduce_class_hist"):
            comm.sum_tensor_(hist)
    global_hist = read_small(hist).numpy().copy()
    if int(global_hist.sum()) != n_total:
        raise GGError(f"score-class table holds {int(global_hist.sum())} candidates, the scoring pass counted {n_total}")
    action, tie_budget, cutoff, tie_total, cut_class = choose_cutoff_class(global_hist, p_max, budget)
    ratio = tie_ratio(tie_budget, tie_total)
    action_d = torch.from_numpy(action).to(dev)
    # ---- 2. what this rank will keep: everything above the cut-off class + its share of the class
    cut_test = ScoreTest(cut_class.numerator, cut_class.denominator, True)   # >= the lowest class that contributes
    have_tie = tie_total > 0
    if dedup:
        # second scoring pass against the cut-off (>=): the bit matrix the chunks are emitted from, and this rank's
        # members of the cut-off class / pairs above it from the class sizes at its row boundaries
        sums = torch.zeros(2, dtype=torch.int64, device=dev)
        row_lo, row_hi = min(b0 * _lib.GG_KF_BLOCK, n), min(b1 * _lib.GG_KF_BLOCK, n)
        extra = _lib.KfDedupExtra(None, 0, 0, sums.data_ptr(), cut_class.numerator, cut_class.denominator, row_lo, row_hi)
        staged.score(cut_test, sq_max, extra)
        local_offsets = staged.count(b0, b1)
        t_lo, t_hi = read_small(sums).tolist()
        local_ties = int(t_lo) - int(t_hi) if have_tie else 0
        local_ge = int(read_small(local_offsets[-1:])[0])
        local_above = local_ge - local_ties
    else:
        lh = local_hist if local_hist is not None else global_hist
        local_ties, local_above = int(lh[action == 2].sum()), int(lh[action == 1].sum())
        local_ge = local_ties + local_above
    ties_per_rank = comm.gather_ints(local_ties) if comm.world > 1 else [local_ties]
    above_per_rank = comm.gather_ints(local_above) if comm.world > 1 else [local_above]
    g_start = sum(ties_per_rank[: comm.rank])
    kept_per_rank = []
    g = 0
    for r in range(comm.world):
        kept_per_rank.append(above_per_rank[r] + ties_admitted(g + ties_per_rank[r], ratio) - ties_admitted(g, ratio))
        g += ties_per_rank[r]
    if g != tie_total or sum(kept_per_rank) != budget:
        raise GGError(f"top-n accounting: {g} members of the cut-off class (table: {tie_total}), "
                      f"{sum(kept_per_rank)} edges kept (budget {budget})")
    mine = kept_per_rank[comm.rank]
    exchange = gather and comm.world > 1
    discard = sink == "discard"
    full_ei = full_w = None
    out = None
    if exchange:
        total, lo = sum(kept_per_rank), sum(kept_per_rank[: comm.rank])
        full_ei = torch.empty((2, max(total, 1)), dtype=torch.int64, device=dev)
        full_w = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        out = (full_ei[:, lo: lo + max(mine, 1)], full_w[lo: lo + max(mine, 1)])
    dest = _SelectSink(dev, mine, g_start, want_w64=comm.world == 1, discard=discard, chunk_cap=chunk_cap, out=out)
    # ---- 3. emission
    if not streamed:
        step = 1 << 30   # gg_kf_select_packed takes at most 2^31 - 1 words per call
        for lo in range(0, staged.n_recs, step):
            dest.select(ordered_all[lo: lo + step], min(step, staged.n_recs - lo), sqnorm, p_max, action_d, ratio)
    else:
        if dedup:
            # the shard's offsets are exact: cut it into runs of row blocks of at most chunk_cap words (one read)
            stride = _row_blocks(n) * _lib.GG_KF_BLOCK
            marks = read_small(local_offsets[:: stride][: b1 - b0 + 1].contiguous()).tolist()
            chunks, lo = [], b0
            while lo < b1:
                hi = lo + 1
                while hi < b1 and marks[hi + 1 - b0] - marks[lo - b0] <= chunk_cap:
                    hi += 1
                chunks.append((lo, hi))
                lo = hi
            todo = list(reversed(chunks))
        else:
            staged.recs = staged.rowcnt = None       # the first pass's records are not used: free them
            todo = list(reversed(_plan_chunks(n, b0, b1, local_ge, chunk_cap)))
        while todo:
            c0, c1 = todo.pop()
            if dedup:
                base, end = int(marks[c0 - b0]), int(marks[c1 - b0])
                found = end - base
            else:
                with _stage_prefix("re_"):
                    res = kf_emit(counts, sqnorm, cut_test, sq_max, c0, c1, impl, rec_capacity=chunk_cap,
                                  hist_m=st.row_sum, retry=False)
                recs, found, rowcnt, params, scored = res
                if res.slots > chunk_cap and c1 - c0 > 1:    # denser than planned: halve the range
                    del recs, rowcnt, res
                    mid = (c0 + c1) // 2
                    todo.extend([(mid, c1), (c0, mid)])
                    continue
            if found == 0:
                continue
            ordered = torch.empty(found, dtype=torch.int64, device=dev)
            if dedup:
                # positions in local_offsets count from the shard's first edge: shift the output pointer instead
                staged.place_range(c0, c1, C.c_void_p(local_offsets.data_ptr() + 8 * (c0 - b0) * stride),
                                   C.c_void_p(ordered.data_ptr() - 8 * base), end)
            else:
                if res.slots > chunk_cap:           # a single row block beyond the chunk size: score it with room
                    with _stage_prefix("re_"):
                        recs, found, rowcnt, params, scored = kf_emit(counts, sqnorm, cut_test, sq_max, c0, c1, impl,
                                                                      rec_capacity=res.slots + res.slots // 16 + (1 << 16),
                                                                      hist_m=st.row_sum)
                StagedRecords(recs, found, rowcnt, params, scored).place(ptr(ordered))
                del recs, rowcnt
            dest.select(ordered, found, sqnorm, p_max, action_d, ratio)
            del ordered
    g_end, kept = dest.totals()
    if kept != mine or g_end != g_start + local_ties:
        raise GGError(f"top-n select kept {kept} edges (expected {mine}) and saw {g_end - g_start} members of the "
                      f"cut-off class (expected {local_ties})")
    if exchange:
        with _stage("c2_exchange_edges"):
            for req in comm.exchange_slices([full_ei[0], full_ei[1], full_w], kept_per_rank):
                req.wait()
        total = sum(kept_per_rank)
        return KFEdges(full_ei[:, :total], None, full_w[:total], n_total, cutoff, tie_budget, tie_total, impl,
                       per_rank=kept_per_rank)
    shown = dest.last if discard else mine
    return KFEdges(dest.ei[:, :shown], None if dest.w64 is None else dest.w64[:shown], dest.w32[:shown], n_total, cutoff,
                   tie_budget, tie_total, impl, per_rank=kept_per_rank, sharded=comm.world > 1, resident=not discard)


# --------------------------------------------------------------------------- whole path on one GPU
# --------------------------------------------------------------------------- f4b per-row kNN (faiss_ann mode)
KNN_METRICS = {"IP": 0, "L2": 1}
_KNN_NO_RANK = 0xFFFF


@lru_cache(maxsize=64)
def knn_tables(sq_values: tuple, metric: str):
    """Host tables of gg_knn_*: ``sq_values`` = the distinct squared norms present (sorted).  Returns
    (sqid uint8[sq_max + 1], rank_tab uint16[n_sq, dot_max + 1], rank_q int32[n_ranks], n_ranks).
    rank 0 = most similar; equal similarity <=> equal rank.  "IP": cosine order of a row = order of dot / sqrt(sq_j),
    compared exactly as the rational dot^2 / sq_j; zero dots are never neighbours (similarity 0 is filtered,
    gnn_utils.py:221).  "L2": order of sq_j - 2 dot (ascending)."""
    sq_max = int(sq_values[-1])
    dot_max = sq_max
    sqid = np.full(sq_max + 1, 0xFF, dtype=np.uint8)
    keys = {}
    for ci, sq in enumerate(sq_values):
        sqid[sq] = ci
        for dot in range(dot_max + 1):
            if dot * dot > sq * sq_max:       # Cauchy-Schwarz: no such pair exists
                continue
            if metric == "IP":
                if dot == 0 or sq == 0:
                    continue
                keys[(ci, dot)] = -Fraction(dot * dot, sq)
            else:
                keys[(ci, dot)] = sq - 2 * dot
    ordered = sorted(set(keys.values()))
    rank_of = {v: r for r, v in enumerate(ordered)}
    tab = np.full((len(sq_values), dot_max + 1), _KNN_NO_RANK, dtype=np.uint16)
    for (ci, dot), v in keys.items():
        tab[ci, dot] = rank_of[v]
    rank_q = np.array([int(v) for v in ordered], dtype=np.int32) if metric == "L2" else np.zeros(max(len(ordered), 1), np.int32)
    return sqid, tab, rank_q, max(len(ordered), 1)


@dataclass
class KnnEdges:
    edge_index: torch.Tensor   # int64 [2, E]: source row, neighbour; grouped by source, most similar first
    weight: torch.Tensor       # float32 [E]
    k: int
    metric: str


KNN_ROW_BLOCK = 128   # gg_knn_* row-shard granularity


def knn_row_shard(n: int, world: int, rank: int):
    """Contiguous, 128-aligned row range of ``rank``: every row costs the same (it meets all n columns)."""
    blocks = (n + KNN_ROW_BLOCK - 1) // KNN_ROW_BLOCK
    lo = blocks * rank // world * KNN_ROW_BLOCK
    hi = min(blocks * (rank + 1) // world * KNN_ROW_BLOCK, n)
    return min(lo, n), hi


KNN_IMPL = {"auto": 0, "mma": 1, "tcgen05": 2, "sparse": 3}


def kf_knn_edges(counts: torch.Tensor, sqnorm: torch.Tensor, k_nn: int, distance: str = "L2", hist_mode: int = -1,
                 rows=None, qmax_reduce=None, comm=None, gather: bool = True, impl: str = "auto") -> KnnEdges:
    """Exact per-row k-nearest-neighbour KF edges: what kf_faiss_to_edge_index_weight (gnn_utils.py:122-237)
    approximates through FAISS.  Two passes over the Gram matrix (gg_knn_count / gg_knn_write), no sort by value.

    Rows are independent, so the work shards by row range: with ``comm`` (dist.TorchComm) every rank takes
    knn_row_shard() of the rows against all columns, the ranks agree on the "L2" normaliser (max over ranks of one
    integer) and, unless ``gather=False``, all-gather their edge lists in rank order = the single-GPU result.
    ``rows=(lo, hi)`` computes one shard explicitly (``qmax_reduce``: callable mapping the shard's integer
    maximum distance to the global one).  ``impl``: "auto" = the sparse inverted-index join for rows of 64..1024 counts
    with at most 16 non-zero entries on up to 200,000 nodes, else both Gram passes on the tcgen05 CTA-pair kernel for
    rows of 128..1024 bytes, else the mma.sync kernel; "sparse" / "tcgen05" / "mma" force one (tests).
    ``hist_mode`` forces where pass 1 of the mma.sync kernel accumulates (tests): 0 global atomics, 1 / 2 shared memory
    16- / 32-bit."""
    if distance not in KNN_METRICS:
        raise ValueError("distance must be 'IP' or 'L2'")
    if impl not in KNN_IMPL:
        raise ValueError("impl must be 'auto', 'mma', 'tcgen05' or 'sparse'")
    lib = _lib.load()
    check(lib.gg_knn_set_impl(KNN_IMPL[impl]), "gg_knn_set_impl")
    dev = counts.device
    n, pitch = int(counts.shape[0]), int(counts.shape[1])
    empty = KnnEdges(torch.zeros((2, 0), dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.float32, device=dev),
                     int(k_nn), distance)
    if k_nn < 1 or n < 2:
        return empty
    if k_nn > n - 1:
        raise ValueError("k must be at most N - 1 (the reference searches k + 1 <= N entries)")
    sqnorm = sqnorm[:n].contiguous()
    sq_values = tuple(int(v) for v in torch.unique(sqnorm).cpu().tolist())
    if sq_values[-1] == 0:
        return empty
    if sq_values[-1] > 254:
        raise GGError("kNN KF path: squared row norms above 254 are outside the k-mer count range")
    sharded = comm is not None and comm.world > 1
    if rows is None:
        rows = knn_row_shard(n, comm.world, comm.rank) if sharded else (0, n)
    lo, hi = int(rows[0]), int(rows[1])
    m = hi - lo
    metric = KNN_METRICS[distance]
    sqid, tab, rank_q, n_ranks = knn_tables(sq_values, distance)
    n_sq, dot_max, sq_max = int(tab.shape[0]), int(tab.shape[1]) - 1, int(sq_values[-1])
    sqid_d = torch.from_numpy(sqid).to(dev)
    tab_d = torch.from_numpy(tab.view(np.int16)).to(dev)
    rank_q_d = torch.from_numpy(rank_q).to(dev)
    st = _stream()
    n_edges, qmax_v = 0, 0
    qmax = torch.zeros(1, dtype=torch.int32, device=dev)
    # rows of sub-k-mer counts are sparse (<= k - small_k + 1 non-zero entries): wide rows go through the inverted-index
    # join, which never forms the zero dots; a forced histogram mode names the mma.sync kernel
    index = None
    # (below 128 columns the posting lists are so long that the join does more work than the dense kernels: 48 against
    # 20 ms at 64 columns -- only on request there)
    if (m > 0 and hist_mode < 0 and (impl == "sparse" or (impl == "auto" and pitch >= 128))
            and lib.gg_knn_sparse_eligible(pitch, n, n_sq, n_ranks)):
        ib = int(lib.gg_knn_sparse_index_bytes(n, pitch))
        index = torch.empty(ib, dtype=torch.uint8, device=dev)
        idx_status = torch.empty(2, dtype=torch.int64, device=dev)
        with _stage(f"knn_index_d{pitch}"):
            check(lib.gg_knn_sparse_index(ptr(counts), pitch, n, ptr(sqnorm), ptr(sqid_d), sq_max, n_sq, ptr(index), ib,
                                          ptr(idx_status), st), "gg_knn_sparse_index")
        if int(read_small(idx_status)[0]) != 0:      # a row with more than 16 non-zero counts: not k-mer count rows
            index = None
    if index is None and impl == "sparse":
        raise GGError("the sparse kNN path cannot serve this shape (columns, nodes, classes or non-zero counts per row)")
    if m > 0:
        hist = torch.empty(m * n_ranks, dtype=torch.int32, device=dev)
        with _stage(f"knn_count_d{pitch}"):
            if index is not None:
                check(lib.gg_knn_sparse_count(ptr(index), ib, pitch, n, lo, hi, ptr(sqnorm), ptr(tab_d), n_sq, dot_max,
                                              n_ranks, ptr(hist), st), "gg_knn_sparse_count")
            else:
                check(lib.gg_knn_count(ptr(counts), pitch, n, lo, hi, ptr(sqnorm), ptr(sqid_d), sq_max, ptr(tab_d), n_sq,
                                       dot_max, n_ranks, int(hist_mode), ptr(hist), st), "gg_knn_count")
        cut = torch.empty(m, dtype=torch.int32, device=dev)
        tie_left = torch.empty(m, dtype=torch.int32, device=dev)
        row_off = torch.empty(m + 1, dtype=torch.int64, device=dev)
        wb = int(lib.gg_knn_plan_workspace_bytes(m))
        ws = torch.empty(wb, dtype=torch.uint8, device=dev)

        def plan(phase):
            with _stage("knn_plan"):
                check(lib.gg_knn_plan(n, lo, hi, n_ranks, int(k_nn), metric, ptr(sqnorm), ptr(rank_q_d), ptr(hist),
                                      ptr(cut), ptr(tie_left), ptr(row_off), ptr(qmax), phase, ptr(ws), wb, st),
                      "gg_knn_plan")
    two_phase = metric == 1 and (sharded or qmax_reduce is not None)
    if two_phase:   # the "L2" normaliser is global: max over the shards of one integer
        if m > 0:
            plan(1)
        mine = int(read_small(qmax.to(torch.int64))[0])
        qmax_v = max(comm.gather_ints(mine)) if sharded else int(qmax_reduce(mine))
        qmax.fill_(qmax_v)
        if m > 0:
            plan(2)
    elif m > 0:
        plan(0)
        if metric == 1:
            qmax_v = int(read_small(qmax.to(torch.int64))[0])
    if metric == 1 and qmax_v == 0:
        raise GGError("kNN KF path (L2): every kept distance is zero, the reference's 1 - d / max(d) is undefined")
    if m > 0:
        n_edges = int(read_small(row_off[m:m + 1])[0])
    edge_index = torch.empty((2, n_edges), dtype=torch.int64, device=dev)
    weight = torch.empty(n_edges, dtype=torch.float32, device=dev)
    if n_edges:
        words = n_edges if index is not None else int(lib.gg_knn_write_scratch_words(pitch, n, n_sq, dot_max, n_ranks, m, n_edges))
        recs = torch.empty(words, dtype=torch.int64, device=dev)
        with _stage(f"knn_write_d{pitch}"):
            if index is not None:
                check(lib.gg_knn_sparse_write(ptr(index), ib, pitch, n, lo, hi, ptr(sqnorm), ptr(tab_d), n_sq, dot_max,
                                              n_ranks, metric, ptr(hist), ptr(cut), ptr(tie_left), ptr(row_off), ptr(qmax),
                                              ptr(recs), C.c_void_p(edge_index.data_ptr()),
                                              C.c_void_p(edge_index.data_ptr() + 8 * n_edges), ptr(weight), n_edges, st),
                      "gg_knn_sparse_write")
            else:
                check(lib.gg_knn_write(ptr(counts), pitch, n, lo, hi, ptr(sqnorm), ptr(sqid_d), sq_max, ptr(tab_d), n_sq,
                                       dot_max, n_ranks, metric, ptr(hist), ptr(cut), ptr(tie_left), ptr(row_off),
                                       ptr(qmax), ptr(recs), words, C.c_void_p(edge_index.data_ptr()),
                                       C.c_void_p(edge_index.data_ptr() + 8 * n_edges), ptr(weight), n_edges, st),
                      "gg_knn_write")
    if sharded and gather:
        with _stage("c4_gather_knn_edges"):
            edge_index = comm.gather_cat(edge_index, dim=1)
            weight = comm.gather_cat(weight, dim=0)
    return KnnEdges(edge_index, weight, int(k_nn), distance)


@dataclass
class BuiltGraph:
    k: int
    db: DBGraph
    kf_counts: List[KFCounts]
    x: List[torch.Tensor]              # float32 [N, D'_i] per small_k
    kf: List[KFEdges] = field(default_factory=list)


def build_graph(sequences, k: int, small_k="2", create_all_kmers=False, edges_threshold=0.0, edges_keep_top_k=None,
                kf_impl="auto", with_kf=True, comm=None, kf_gather: bool = True, kf_sink: str = "device",
                kf_chunk_records: Optional[int] = None, keep_features: bool = True) -> BuiltGraph:
    """kmer_ssgnn.py:162-213 + sampling_task.py:128-161.

    Single GPU by default.  With ``comm`` (dist.TorchComm) ``sequences`` is this
    rank's shard of the reads (a ReadStream whose pos_base is the shard's global
    stream offset): histograms are summed and first occurrences min-reduced over
    ranks, the De Bruijn assembly and features run replicated, KF is sharded by
    row block."""
    comm = comm or LocalComm()
    if hasattr(comm, "begin_build"):
        comm.begin_build()
    stream = upload_reads(sequences)
    fused_kcs = None
    if comm.world > 1:
        from . import dist

        capacity = 1 << 16
        while True:  # K1 without a host round trip; reduce_counts reads every rank's status in its one device read
            try:
                piece = count_piece(k, stream.data.device, comm)
                local = count_kp1mers(stream, k, bridge_capacity=capacity, sync=False, piece=piece)
                with _stage("c1_reduce_counts"):
                    cr = dist.reduce_counts(local, comm, piece=piece)
                break
            except OverflowError as exc:  # same decision on every rank: the counts were all-reduced
                capacity = int(exc.args[0])
        db = build_db(cr, normalise=True, create_all_kmers=create_all_kmers)
    else:
        # K1, K2 and K3's counts back to back with one device read for all three (count_db_counts); streams with
        # bridge records (internal N runs) take K1 and K2 with one read (build_db), and the rare stream with more of them
        # than the default room is counted again with room for all
        db, fused_kcs, cr = count_db_counts(stream, k, [int(s) for s in str(small_k).split(",")], create_all_kmers)
        if db is None:
            try:   # (a count whose status is still unread, n_bridges == -1, is assembled speculatively once more)
                db = build_db(cr, normalise=True, create_all_kmers=create_all_kmers)
            except _NeedBridges as need:
                cr = count_kp1mers(stream, k, bridge_capacity=int(need.args[0]))
                db = build_db(cr, normalise=True, create_all_kmers=create_all_kmers)
    if db.num_edges == 0:
        raise ValueError("no k-mer pairs in the input (the reference raises IndexError at utils.py:122)")
    small = [int(s) for s in str(small_k).split(",")]
    kcs = fused_kcs if fused_kcs is not None else kf_counts_many(db.node_codes, k, small)
    if with_kf and comm.world == 1 and kf_gather and kf_sink == "device" and kf_chunk_records is None:
        # one GPU: the float features are queued between the placements, where they cover a size read
        out = BuiltGraph(k, db, kcs, [])

        def features():
            if keep_features:
                out.x = [kf_features_f32(kc) for kc in kcs]

        out.kf = kf_sets_overlapped(kcs, edges_threshold, edges_keep_top_k, kf_impl, between=features)
        return out
    xs = [kf_features_f32(kc) for kc in kcs] if keep_features else []
    out = BuiltGraph(k, db, kcs, xs)
    if with_kf:
        b0, b1 = 0, None
        if comm.world > 1:
            from . import dist

            b0, b1 = dist.kf_block_range(_row_blocks(db.num_nodes), comm.world, comm.rank)
        # every set is scored first; ONE exchange / device read then brings the candidate counts of all of them.  Wide
        # sets are queued first and the narrow (duplicate-row) sets only start their class kernels: the host then waits
        # for the class count while the tensor kernel of the wide set runs (kf_stage(defer=True))
        stages = [None] * len(kcs)
        for i in sorted(range(len(kcs)), key=lambda j: -int(kcs[j].counts.shape[1])):
            kc = kcs[i]
            stages[i] = kf_stage(kc.counts, kc.sqnorm, edges_threshold, edges_keep_top_k, sq_max=kc.m * kc.m, impl=kf_impl,
                                 iblock_begin=b0, iblock_end=b1, row_sum=kc.m, defer=True)
        for st in stages:
            if st.finish_stage is not None:
                st.finish_stage()
        kf_resolve(stages, comm)
        out.kf = [kf_finish(st, comm, gather=kf_gather, chunk_records=kf_chunk_records, sink=kf_sink) for st in stages]
        for kf in out.kf:  # multi-GPU: the edge exchanges of earlier sets ran underneath the scoring of later ones
            kf.wait()
    return out


# --------------------------------------------------------------------------- host-to-host build with overlapped copies
@dataclass
class HostGraph:
    """Every tensor the reference's builders hand to PyG, in pinned host memory."""
    k: int
    num_nodes: int
    node_codes: torch.Tensor
    edge_index: torch.Tensor
    weight: torch.Tensor
    x: List[torch.Tensor]
    kf_edge_index: List[torch.Tensor]
    kf_weight: List[torch.Tensor]
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    db: Optional["DBGraph"] = None                 # the device-resident originals (builders keep them for transform())
    kf_counts: Optional[List["KFCounts"]] = None
    marshal_ms: float = 0.0                        # host time spent turning list[str] input into the byte stream


def graph_to_host(built: "BuiltGraph", host_expand: bool = True) -> "HostGraph":
    """Every tensor of a device-resident graph in pinned host memory (what the CPU-side loaders want).  Large
    transfers first; wide feature matrices cross PCIe as their uint8 counts and are expanded by host threads while the
    edge lists follow (gg_host_features_f32), exactly as in build_graph_to_host."""
    lib = _lib.load()
    dev = built.db.node_codes.device
    d2h = 0

    def down(t):
        nonlocal d2h
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        d2h += t.numel() * t.element_size()
        return h

    jobs, h_x = [], []
    for kc, x in zip(built.kf_counts, built.x):
        if not (host_expand and kc.counts.shape[1] >= 256 and int(kc.col_keep.numel()) == kc.counts.shape[1]):
            h_x.append(down(x))
            continue
        h_counts = down(kc.counts)
        landed = torch.cuda.Event()
        landed.record()
        out = torch.empty(kc.counts.shape, dtype=torch.float32, pin_memory=True)
        lut = value_lut(kc.m).astype(np.float32)
        job = lib.gg_host_features_f32_start(C.c_void_p(h_counts.data_ptr()), h_counts.numel(),
                                             C.c_void_p(lut.ctypes.data), C.c_void_p(out.data_ptr()), _HOST_THREADS,
                                             dev.index, C.c_void_p(landed.cuda_event))
        if not job:
            raise GGError("gg_host_features_f32_start failed: " + lib.gg_last_error().decode("utf-8", "replace"))
        jobs.append((job, h_counts, out, lut, landed))
        h_x.append(out)
    try:
        h_codes, h_ei, h_w = down(built.db.node_codes), down(built.db.edge_index), down(built.db.weight)
        kf_ei = [down(kf.edge_index) for kf in built.kf]
        kf_w = [down(kf.weight) for kf in built.kf]
        torch.cuda.current_stream().synchronize()
    finally:
        rcs = [lib.gg_host_job_wait(C.c_void_p(job[0])) for job in jobs]
    for rc in rcs:
        check(rc, "gg_host_job_wait")
    return HostGraph(built.k, built.db.num_nodes, h_codes, h_ei, h_w, h_x, kf_ei, kf_w, 0, d2h)


_COPY_STREAMS = {}


def _copy_stream():
    dev = torch.cuda.current_device()
    if dev not in _COPY_STREAMS:
        _COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[dev]


# measured on the 16-core GPU box: 8 threads 14.7 ms, 16 threads 15.1 ms (they fight the DMA for memory bandwidth), 4 threads 17.1 ms
_HOST_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))


class _PackedReads:
    """list[str] -> ASCII stream in pinned memory through libgg_pyhost.so (csrc/gg_pyhost.c): one pass over the list
    for (pointer, length) of every read, then native threads copy read ranges while the GIL is released."""

    def __init__(self, reads, threads):
        self.lib = _lib.load_pyhost()
        info = (C.c_int64 * 4)()
        self.handle = self.lib.gg_py_reads_scan(reads, threads, info)
        if not self.handle:
            if info[3] >= 0:
                raise ValueError(f"read {int(info[3])} is not an ASCII str")
            raise ValueError("reads must be a list or tuple of str")
        self.n_reads, self.n_bytes, self.fixed_len = int(info[0]), int(info[1]), int(info[2])
        self.threads = threads
        self.flat = torch.empty(max(self.n_bytes, 1), dtype=torch.uint8, pin_memory=True)

    def copy(self, lo, hi):
        if self.lib.gg_py_reads_copy(self.handle, lo, hi, C.c_void_p(self.flat.data_ptr()), self.threads) != 0:
            raise GGError("gg_py_reads_copy rejected its arguments")

    def close(self):
        if self.handle:
            self.lib.gg_py_reads_free(self.handle)
            self.handle = None


def build_graph_to_host(reads, k: int, small_k="2", create_all_kmers=False, edges_threshold=0.0,
                        edges_keep_top_k=None, kf_impl="auto", chunks: int = 8, host_expand: bool = True,
                        with_kf: bool = True) -> HostGraph:
    """The whole path from host reads to host graph tensors (what ``create_graphs`` hands to the CPU-side PyG
    loaders), with the PCIe copies overlapped with the kernels: the reads go up in ``chunks`` pieces that are
    counted as they land, the DB graph and the features go down while the KF sets are scored, and each KF set
    goes down while the next one is scored.

    ``reads``: uint8 [n_reads, read_len] ASCII on the host (pinned memory avoids a staging copy), or the reference's
    own input type, a list of read strings (kmer_ssgnn.py:56-57) -- marshalled natively, piece by piece, under the
    uploads (ragged reads go up in one piece)."""
    _require_cuda()
    lib = _lib.load()
    if not (1 <= k <= _lib.GG_MAX_K):
        raise ValueError(f"k must be in [1, {_lib.GG_MAX_K}]")
    packed, marshal_ms = None, 0.0
    if isinstance(reads, (list, tuple)):
        import time as _time

        t0 = _time.perf_counter()
        packed = _PackedReads(reads, _HOST_THREADS)
        marshal_ms = (_time.perf_counter() - t0) * 1e3
        n_reads, length, n_bytes, flat = packed.n_reads, packed.fixed_len, packed.n_bytes, packed.flat
        if n_reads == 0:
            packed.close()
            raise ValueError("no k-mer pairs in the input (the reference raises IndexError at utils.py:122)")
    elif isinstance(reads, torch.Tensor) and reads.dtype == torch.uint8 and reads.dim() == 2 and not reads.is_cuda:
        n_reads, length = int(reads.shape[0]), int(reads.shape[1])
        n_bytes = n_reads * length
        flat = reads.reshape(-1)
    else:
        raise ValueError("reads must be a host uint8 [n_reads, read_len] tensor or a list of str")
    dev = torch.device("cuda", torch.cuda.current_device())
    main, side = torch.cuda.current_stream(), _copy_stream()
    buf = torch.empty(_padded(n_bytes), dtype=torch.uint8, device=dev)
    bins = int(lib.gg_hist_bins(k))
    hist = torch.empty(bins, dtype=torch.int32, device=dev)
    first = torch.empty(bins, dtype=torch.int64, device=dev)
    status = torch.empty(_lib.ST_WORDS, dtype=torch.int64, device=dev)
    cap = 1 << 16
    bridges = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    st = _stream()
    check(lib.gg_count_reset(k, ptr(hist), ptr(first), ptr(status), st), "gg_count_reset")
    side.wait_stream(main)
    buf.record_stream(side)
    if length > 0:   # fixed-length rows: pieces of whole reads, a multiple of 16 of them so that every piece is aligned
        per = max(16, -(-n_reads // max(chunks, 1) // 16) * 16)
        pieces = [(lo, min(n_reads, lo + per), lo * length, min(n_reads, lo + per) * length) for lo in range(0, n_reads, per)]
    else:            # '\n'-separated stream: one piece
        per = n_reads
        pieces = [(0, n_reads, 0, n_bytes)]
    piece_bytes = max(b1 - b0 for _, _, b0, b1 in pieces)
    ws_bytes = int(lib.gg_count_part_workspace_bytes(piece_bytes, k))  # one workspace serves every piece in turn
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    try:
        for lo, hi, b0, b1 in pieces:
            if packed is not None:
                import time as _time

                t0 = _time.perf_counter()
                packed.copy(lo, hi)       # native threads, GIL released; earlier pieces are uploading / being counted
                marshal_ms += (_time.perf_counter() - t0) * 1e3
            with torch.cuda.stream(side):
                buf[b0:b1].copy_(flat[b0:b1], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(side)
            main.wait_event(landed)
            piece = C.c_void_p(buf.data_ptr() + b0)
            with _stage("k1_count"):
                if ws is not None:
                    check(lib.gg_count_kp1mers_part(piece, b1 - b0, length, k, b0, ptr(hist), ptr(bridges), cap,
                                                    ptr(status), ptr(ws), ws_bytes, st), "gg_count_kp1mers_part")
                else:
                    check(lib.gg_count_kp1mers(piece, b1 - b0, length, k, b0, ptr(hist), ptr(bridges), cap, ptr(status),
                                               st), "gg_count_kp1mers")
    finally:
        if packed is not None:
            side.synchronize()            # the uploads read the staging buffer
            packed.close()
    with _stage("k1_first"):
        check(lib.gg_first_occurrence(ptr(buf), n_bytes, length, k, 0, ptr(hist), ptr(first), ptr(status), st),
              "gg_first_occurrence")
    st_host = read_small(status).tolist()
    if int(st_host[_lib.ST_BAD_BYTES]) > 0:
        raise ValueError(f"{int(st_host[_lib.ST_BAD_BYTES])} input bytes outside the alphabet ACGTN")
    nb = int(st_host[_lib.ST_BRIDGES])
    if nb > cap:  # many internal N runs: redo the count in one piece with room for every bridge record
        cr = count_kp1mers(ReadStream(buf, n_bytes, length, n_reads), k, bridge_capacity=nb)
    else:
        cr = CountResult(k, hist, first, bridges, nb, status, int(st_host[_lib.ST_NONZERO]), pos_limit=n_bytes)
    if cr.n_nonzero == 0 and cr.n_bridges == 0:
        raise ValueError("no k-mer pairs in the input (the reference raises IndexError at utils.py:122)")
    db = build_db(cr, normalise=True, create_all_kmers=create_all_kmers)
    kcs = kf_counts_many(db.node_codes, k, [int(s) for s in str(small_k).split(",")])
    # Wide feature matrices cross PCIe as their uint8 counts and are expanded by the host cores while the edge lists are
    # still on the wire (gg_host_features_f32); narrow ones, and matrices with dropped columns, come down as float32.
    via_host = [host_expand and kc.counts.shape[1] >= 256 and int(kc.col_keep.numel()) == kc.counts.shape[1] for kc in kcs]
    xs = [None if h else kf_features_f32(kc) for kc, h in zip(kcs, via_host)]

    d2h = 0
    keep_alive = []

    def down(tensors):
        """Queue device->pinned-host copies on the side stream, ordered after what the main stream has produced."""
        nonlocal d2h
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        out = []
        with torch.cuda.stream(side):
            for t in tensors:
                t.record_stream(side)
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t, non_blocking=True)
                out.append(h)
                d2h += t.numel() * t.element_size()
        keep_alive.extend(tensors)
        return out

    h_codes, h_ei, h_w, *h_small = down([db.node_codes, db.edge_index, db.weight, *[x for x in xs if x is not None]])
    h_small = iter(h_small)
    h_x, workers = [], []
    for kc, h in zip(kcs, via_host):
        if not h:
            h_x.append(next(h_small))
            continue
        (h_counts,) = down([kc.counts])
        landed = torch.cuda.Event()
        landed.record(side)
        out = torch.empty(kc.counts.shape, dtype=torch.float32, pin_memory=True)
        lut = value_lut(kc.m).astype(np.float32)

        # a native thread waits for the event and expands (no interpreter scheduling on the path)
        job = lib.gg_host_features_f32_start(C.c_void_p(h_counts.data_ptr()), h_counts.numel(),
                                             C.c_void_p(lut.ctypes.data), C.c_void_p(out.data_ptr()), _HOST_THREADS,
                                             dev.index, C.c_void_p(landed.cuda_event))
        if not job:
            raise GGError("gg_host_features_f32_start failed: " + lib.gg_last_error().decode("utf-8", "replace"))
        workers.append((job, h_counts, out, lut, landed))
        h_x.append(out)
    kf_ei, kf_w = [], []
    try:
        # one set at a time here (not stage-all-first as in build_graph): the first set's edges, the largest transfer
        # of the step, start going down while the second set is still being scored -- measured 17.2 vs 18.5 ms
        for kc in (kcs if with_kf else []):
            kf = kf_edges(kc.counts, kc.sqnorm, edges_threshold, edges_keep_top_k, sq_max=kc.m * kc.m, impl=kf_impl,
                          row_sum=kc.m)
            e, w = down([kf.edge_index, kf.weight])
            kf_ei.append(e)
            kf_w.append(w)
        side.synchronize()
    finally:  # the native expansion jobs hold pointers into h_counts / out: always join them
        rcs = [lib.gg_host_job_wait(C.c_void_p(job[0])) for job in workers]
    for rc in rcs:
        check(rc, "gg_host_job_wait")
    return HostGraph(k, db.num_nodes, h_codes, h_ei, h_w, h_x, kf_ei, kf_w, n_bytes, d2h, db=db, kf_counts=kcs,
                     marshal_ms=marshal_ms)
