"""The reference's graph-construction functions, same names and signatures,
computed by the CUDA library.

    weighted_directed_edges            src/utils.py:52-84
    build_deBruijn_graph               src/utils.py:87-165
    node_embedding_initial             src/representations/gnn_common/gnn_utils.py:428-493
    cosine_similarity_to_edge_index_weight                        gnn_utils.py:240-322

The reference returns Python dicts / a networkx DiGraph / dense float64 arrays.
The objects returned here behave like those (same iteration order, same
values) but keep their payload on the GPU and only materialise host-side
Python structures when a caller actually iterates them, so that
``create_graphs`` chains the four calls without a host round trip.
"""
from collections.abc import Mapping
from typing import Optional

import numpy as np
import torch

from . import core
from ._lib import GGError

ALPHABET = ["A", "C", "G", "T"]
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def decode_kmers(codes, k):
    codes = np.asarray(codes, dtype=np.int64)
    out = np.empty((codes.shape[0], k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = _BASES[(codes >> (2 * (k - 1 - j))) & 3]
    return [row.tobytes().decode("ascii") for row in out]


def encode_kmers(kmers, k):
    arr = np.frombuffer("".join(kmers).encode("ascii"), dtype=np.uint8).reshape(-1, k)
    lut = np.full(256, -1, dtype=np.int64)
    lut[_BASES] = np.arange(4)
    digits = lut[arr]
    if (digits < 0).any():
        raise ValueError("k-mer with a character outside ACGT")
    codes = np.zeros(arr.shape[0], dtype=np.int64)
    for j in range(k):
        codes = codes * 4 + digits[:, j]
    return codes


# --------------------------------------------------------------------------- a2
class _LazyDict(Mapping):
    """Read-only insertion-ordered mapping whose dict is built on first use."""

    def _build(self):  # pragma: no cover - abstract
        raise NotImplementedError

    def _d(self):
        if getattr(self, "_dict", None) is None:
            self._dict = self._build()
        return self._dict

    def __getitem__(self, key):
        return self._d()[key]

    def __iter__(self):
        return iter(self._d())

    def __len__(self):
        return len(self._d())

    def __repr__(self):
        return f"<{type(self).__name__} k={self.k} (device-resident)>"


class PairFrequency(_LazyDict):
    """pair_frequency of weighted_directed_edges: {(kmer_i, kmer_next): count} in
    first-occurrence order (src/utils.py:82)."""

    def __init__(self, count_result: core.CountResult):
        self.count_result = count_result
        self.k = count_result.k
        self._dict = None

    def arrays(self):
        """(u_codes, v_codes, counts, first_pos) int64 host arrays in dict order."""
        cr = self.count_result
        k = cr.k
        hist = cr.hist.cpu().numpy().view(np.uint32)
        first = cr.first_pos.cpu().numpy().view(np.uint64)
        nz = np.nonzero(hist)[0]
        u, v = nz >> 2, nz & (4**k - 1)
        cnt, pos = hist[nz].astype(np.int64), first[nz]
        if cr.n_bridges:
            br = cr.bridges[: cr.n_bridges].cpu().numpy()
            bu, bv = br[:, 0].view(np.uint32).astype(np.int64), br[:, 1].view(np.uint32).astype(np.int64)
            bpos = br[:, 2:4].copy().view(np.uint64).reshape(-1)
            key = bu << (2 * k) | bv
            uniq, inv, bc = np.unique(key, return_inverse=True, return_counts=True)
            bmin = np.full(uniq.shape[0], np.iinfo(np.uint64).max, dtype=np.uint64)
            np.minimum.at(bmin, inv, bpos)
            u = np.concatenate([u, uniq >> (2 * k)])
            v = np.concatenate([v, uniq & (4**k - 1)])
            cnt = np.concatenate([cnt, bc.astype(np.int64)])
            pos = np.concatenate([pos, bmin])
        order = np.argsort(pos, kind="stable")
        return u[order].astype(np.int64), v[order].astype(np.int64), cnt[order], pos[order].astype(np.int64)

    def _build(self):
        u, v, cnt, _ = self.arrays()
        ks, vs = decode_kmers(u, self.k), decode_kmers(v, self.k)
        return {(a, b): int(c) for a, b, c in zip(ks, vs, cnt)}


class ItemFrequency(_LazyDict):
    """item_frequency of weighted_directed_edges: {kmer: times it is the source of a pair}."""

    def __init__(self, pairs: PairFrequency):
        self.pairs = pairs
        self.k = pairs.k
        self._dict = None

    def _build(self):
        out = {}
        for (a, _), c in self.pairs.items():
            out[a] = out.get(a, 0) + c
        return out


def weighted_directed_edges(sequences, k: int = 3, stride: int = 1, inlcudeUNK: bool = True,
                            disable_tqdm: bool = True):
    """src/utils.py:52-84.  The GPU path covers what the builders call it with:
    stride 1, inlcudeUNK=False (kmer_ssgnn.py:174-180)."""
    if inlcudeUNK:
        raise NotImplementedError("inlcudeUNK=True keeps an 'N'*k node; the builders always pass False")
    if stride != 1:  # higher-stride graphs (kmer_ssgnn.py:215-226): pairs are no longer (k+1)-mers
        pairs = PairFrequency(core.count_pairs_stride(sequences, k, int(stride)))
        return ItemFrequency(pairs), pairs
    stream = core.upload_reads(sequences)
    pairs = PairFrequency(core.count_kp1mers(stream, k))
    return ItemFrequency(pairs), pairs


# --------------------------------------------------------------------------- a3 / a4
class KmerGraph:
    """What build_deBruijn_graph returns, in the shape the callers use it:
    ``nodes()`` in insertion order (kmer_ssgnn.py:312-314), ``number_of_nodes()``,
    ``edges(data=...)``; plus the from_networkx tensors, already on the device."""

    def __init__(self, db: core.DBGraph):
        self.db = db
        self.k = db.k
        self._nodes = None

    # networkx-like surface
    def nodes(self):
        if self._nodes is None:
            self._nodes = decode_kmers(self.db.node_codes.cpu().numpy(), self.k)
        return list(self._nodes)

    def number_of_nodes(self):
        return self.db.num_nodes

    def number_of_edges(self):
        return self.db.num_edges

    def __len__(self):
        return self.db.num_nodes

    def __iter__(self):
        return iter(self.nodes())

    def __contains__(self, kmer):
        return kmer in set(self.nodes())

    def edges(self, data=False):
        names = self.nodes()
        ei = self.db.edge_index.cpu().numpy()
        w = self.db.weight64.cpu().numpy()
        for i in range(ei.shape[1]):
            u, v = names[ei[0, i]], names[ei[1, i]]
            if data is False:
                yield (u, v)
            elif data is True:
                yield (u, v, {"weight": float(w[i])})
            else:
                yield (u, v, float(w[i]))

    def to_networkx(self):
        import networkx as nx

        g = nx.DiGraph()
        g.add_nodes_from(self.nodes())
        g.add_weighted_edges_from(self.edges(data="weight"))
        return g

    # from_networkx surface (PyG Data fields)
    @property
    def edge_index(self):
        return self.db.edge_index

    @property
    def weight(self):
        return self.db.weight

    @property
    def num_nodes(self):
        return self.db.num_nodes


def _count_result_from_dict(pair_frequency, device=None):
    """A plain {(kmer, kmer): count} dict (e.g. made by the reference's own
    weighted_directed_edges) -> device histogram in insertion order."""
    core._require_cuda()
    keys = list(pair_frequency.keys())
    if not keys:
        raise IndexError("list index out of range")  # utils.py:122 on an empty dict
    k = len(keys[0][0])
    u = encode_kmers([a for a, _ in keys], k)
    v = encode_kmers([b for _, b in keys], k)
    cnt = np.fromiter((pair_frequency[p] for p in keys), dtype=np.int64, count=len(keys))
    overlap = (u & (4 ** (k - 1) - 1)) == (v >> 2)
    bins = 4 ** (k + 1)
    hist = np.zeros(bins, dtype=np.uint32)
    first = np.full(bins, np.iinfo(np.uint64).max, dtype=np.uint64)
    pos = np.arange(len(keys), dtype=np.uint64)
    e = (u[overlap] << 2) | (v[overlap] & 3)
    hist[e] = cnt[overlap]
    first[e] = pos[overlap]
    nb_rows = []
    for uu, vv, cc, pp in zip(u[~overlap], v[~overlap], cnt[~overlap], pos[~overlap]):
        nb_rows.extend([[uu, vv, pp & 0xFFFFFFFF, pp >> 32]] * int(cc))
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    bridges = torch.tensor(nb_rows if nb_rows else [[0, 0, 0, 0]], dtype=torch.int64).to(torch.int32).to(dev)
    return core.CountResult(k, torch.from_numpy(hist.view(np.int32)).to(dev),
                            torch.from_numpy(first.view(np.int64)).to(dev), bridges, len(nb_rows),
                            torch.zeros(8, dtype=torch.int64, device=dev), int(overlap.sum()),
                            pos_limit=max(len(keys), 1))  # positions are dict indices: bounds K2's sort keys


def build_deBruijn_graph(pair_frequency, normalise: bool = False, remove_N: bool = False,
                         edge_weight_threshold: Optional[float] = None, keep_top_k: Optional[float] = None,
                         create_all_kmers: bool = True, disable_tqdm: bool = True) -> KmerGraph:
    """src/utils.py:87-165.  Unlike the reference this does not mutate ``pair_frequency``.

    Only ``remove_N=True`` is built (what every builder passes, kmer_ssgnn.py:181-187): with ``remove_N=False`` the
    reference keeps an isolated 'N'*k node (utils.py:131,162-163) that shifts every node index, so that case raises
    instead of silently returning a different graph (same policy as ``inlcudeUNK=True`` above)."""
    if not remove_N:
        raise NotImplementedError("remove_N=False keeps an 'N'*k node; the builders always pass remove_N=True")
    if isinstance(pair_frequency, PairFrequency):
        cr = pair_frequency.count_result
    else:
        cr = _count_result_from_dict(pair_frequency)
    if cr.n_nonzero == 0 and cr.n_bridges == 0:
        raise IndexError("list index out of range")  # what utils.py:122 raises on an empty dict
    return KmerGraph(core.build_db(cr, normalise=normalise, create_all_kmers=create_all_kmers,
                                   edge_weight_threshold=edge_weight_threshold, keep_top_k=keep_top_k))


# --------------------------------------------------------------------------- a5
class KFLabels:
    """Dense float64 [N, D'] feature matrix of node_embedding_initial, kept as
    uint8 counts on the device.  Converts to the reference's ndarray on demand
    (np.asarray / np.vstack / indexing)."""

    def __init__(self, kc: core.KFCounts):
        self.kc = kc
        self._host = None

    @property
    def shape(self):
        return (int(self.kc.counts.shape[0]), int(self.kc.col_keep.numel()))

    dtype = np.dtype(np.float64)
    ndim = 2

    def __len__(self):
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            c = self.kc.counts[:, self.kc.col_keep.long()].cpu().numpy()
            self._host = core.value_lut(self.kc.m)[c]
        return self._host if dtype is None else self._host.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.__array__()[idx]

    def __iter__(self):
        return iter(self.__array__())

    def to_torch_float32(self):
        return core.kf_features_f32(self.kc)


def node_embedding_initial(graph, method: str = "onehot", k: int = 3, small_k: Optional[int] = None,
                           random_size: int = 4):
    """gnn_utils.py:428-493."""
    n = graph.number_of_nodes()
    if method == "onehot":
        eye = np.eye(n)
        return [eye[i] for i in range(n)]
    if method == "random":
        return np.random.rand(n, random_size)
    if method != "kmer_frequency":
        raise ValueError("Invalid method for node embedding initialisation")
    if small_k is None:
        raise ValueError("small_k must be provided for kmer_frequency method")
    if not isinstance(graph, KmerGraph):
        codes = torch.from_numpy(encode_kmers(list(graph.nodes()), k)).cuda()
    else:
        codes = graph.db.node_codes
        k = graph.k
    return KFLabels(core.kf_counts(codes, k, int(small_k)))


# --------------------------------------------------------------------------- a6
def _counts_from_matrix(X):
    """Recover integer count rows from a float matrix c * (1/m) (what
    node_embedding_initial produces).  Cosine is scale invariant, so any common
    integer scaling of the rows is as good as the original counts."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be [N, D]")
    pos = X[X > 0]
    if pos.size == 0:
        return np.zeros(X.shape, dtype=np.uint8), 0
    for m in range(1, 65):
        scaled = X * m
        r = np.rint(scaled)
        if np.all(np.abs(scaled - r) <= 1e-9 * np.maximum(1.0, np.abs(r))) and r.max() <= 15 and r.min() >= 0:
            sq = (r.astype(np.int64) ** 2).sum(axis=1)
            if sq.max() <= 225:
                return r.astype(np.uint8), int(sq.max())
    raise NotImplementedError("X is not a small-integer count matrix scaled by a constant; the CUDA KF path "
                              "is exact-integer and only covers node_embedding_initial(kmer_frequency) features")


def _pad_columns(counts_np):
    d = counts_np.shape[1]
    target = next(t for t in (4, 16, 64) if d <= t) if d <= 64 else (d + 127) // 128 * 128
    if target == d:
        return counts_np
    out = np.zeros((counts_np.shape[0], target), dtype=np.uint8)
    out[:, :d] = counts_np
    return out


def cosine_similarity_to_edge_index_weight(X, chunk_size: int = 1024, threshold: float = 0.0,
                                           keep_top_k: Optional[float] = None, impl: str = "auto"):
    """gnn_utils.py:240-322 -> (int64 ndarray [2, E], float64 ndarray [E]).

    Edges come out in the reference's emission order (block pairs of 1024 rows,
    row-major inside).  When the global top-n budget binds, everything strictly
    above the cut-off score is kept plus the first members of the cut-off class
    in that order (the reference's pick inside the class is an artefact of
    np.argpartition)."""
    if chunk_size != 1024:
        raise NotImplementedError("edge order is defined by chunk_size=1024, the only value the reference uses")
    core._require_cuda()
    if isinstance(X, KFLabels):
        kc = X.kc
        res = core.kf_edges(kc.counts, kc.sqnorm, float(threshold), keep_top_k, sq_max=kc.m * kc.m, impl=impl)
    else:
        counts_np, sq_max = _counts_from_matrix(X)
        if sq_max == 0 or counts_np.shape[0] < 2:
            return np.zeros((2, 0), dtype=np.int64), np.zeros(0, dtype=np.float64)
        counts_np = _pad_columns(counts_np)
        counts = torch.from_numpy(counts_np).cuda()
        sqnorm = torch.from_numpy((counts_np.astype(np.int64) ** 2).sum(axis=1).astype(np.int32)).cuda()
        res = core.kf_edges(counts, sqnorm, float(threshold), keep_top_k, sq_max=sq_max, impl=impl)
    return res.edge_index.cpu().numpy(), res.weight64.cpu().numpy()


def kf_faiss_to_edge_index_weight(X, k: int, index_type: str = "IVFFlat", distance: str = "L2", nlist=None, nprobe=None,
                                  m=None, nbits=None, useFloat16: bool = True):
    """gnn_utils.py:122-237 -> (int64 ndarray [2, E], float32 ndarray [E]): every node's k nearest nodes.

    The reference searches an approximate FAISS index; this is the exact search that index approximates, so
    ``index_type`` and its tuning knobs are accepted and ignored.  Edges are grouped by source node, most similar
    first (the order of ``index.search``), ties by ascending target; the query node itself is excluded explicitly
    (the reference drops column 0 of the result, which with duplicate rows may be a duplicate instead)."""
    del index_type, nlist, nprobe, m, nbits, useFloat16
    core._require_cuda()
    if isinstance(X, KFLabels):
        counts, sqnorm = X.kc.counts, X.kc.sqnorm
    else:
        counts_np, sq_max = _counts_from_matrix(X)
        if sq_max == 0 or counts_np.shape[0] < 2:
            return np.zeros((2, 0), dtype=np.int64), np.zeros(0, dtype=np.float32)
        counts_np = _pad_columns(counts_np)
        counts = torch.from_numpy(counts_np).cuda()
        sqnorm = torch.from_numpy((counts_np.astype(np.int64) ** 2).sum(axis=1).astype(np.int32)).cuda()
    res = core.kf_knn_edges(counts, sqnorm, int(k), distance)
    return res.edge_index.cpu().numpy(), res.weight.cpu().numpy()


__all__ = [
    "weighted_directed_edges", "build_deBruijn_graph", "node_embedding_initial",
    "cosine_similarity_to_edge_index_weight", "kf_faiss_to_edge_index_weight", "PairFrequency", "ItemFrequency", "KmerGraph", "KFLabels", "GGError",
]
