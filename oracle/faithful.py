"""Faithful CPU restatement of the reference hot path (test infrastructure only).

Same algorithms, same data structures in spirit (insertion-ordered dicts, a
blocked float64 Gram), written from the behaviour of the reference so that
(i) its outputs can be pinned against the reference's own functions
(tests/golden, oracle/make_golden.py) and (ii) it can be timed on the GPU box's
host cores as the CPU baseline, where /root/reference does not exist.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from collections import defaultdict
from itertools import product

import numpy as np

ALPHABET = "ACGT"


# --------------------------------------------------------------------------- a1
def seq_to_kmers(seq, k=3, stride=1, inlcudeUNK=True):
    """src/utils.py:15-31 -- windows seq[i:i+k] for i = 0, stride, ... with
    i+k <= len(seq); windows holding an 'N' become 'N'*k (inlcudeUNK) or are
    dropped."""
    windows = [seq[i : i + k] for i in range(0, len(seq) - k + 1, stride)]
    if inlcudeUNK:
        pad = "N" * k
        return [pad if "N" in w else w for w in windows]
    return [w for w in windows if "N" not in w]


# --------------------------------------------------------------------------- a2
def weighted_directed_edges(sequences, k=3, stride=1, inlcudeUNK=True, disable_tqdm=True):
    """src/utils.py:52-84 -- counts of consecutive surviving k-mers per read.
    Dict insertion order is first occurrence in the read stream."""
    pair_frequency = defaultdict(int)
    item_frequency = defaultdict(int)
    for seq in sequences:
        kmers = seq_to_kmers(seq, k=k, stride=stride, inlcudeUNK=inlcudeUNK)
        for a, b in zip(kmers, kmers[1:]):
            item_frequency[a] += 1
            pair_frequency[(a, b)] += 1
    return item_frequency, pair_frequency


# --------------------------------------------------------------------------- graph container
class OrderedDiGraph:
    """The slice of networkx.DiGraph behaviour the hot path relies on
    (src/utils.py:120-163): insertion-ordered nodes, per-source insertion-ordered
    successors, remove_node keeps the order of what is left."""

    def __init__(self):
        self._succ = {}

    def add_node(self, n):
        if n not in self._succ:
            self._succ[n] = {}

    def add_edge(self, u, v, weight):
        self.add_node(u)
        self.add_node(v)
        self._succ[u][v] = weight

    def remove_edge(self, u, v):
        del self._succ[u][v]

    def remove_node(self, n):
        del self._succ[n]
        for nbrs in self._succ.values():
            nbrs.pop(n, None)

    def nodes(self):
        return list(self._succ.keys())

    def number_of_nodes(self):
        return len(self._succ)

    def edges(self, data=False):
        for u, nbrs in self._succ.items():
            for v, w in nbrs.items():
                yield (u, v, w) if data else (u, v)

    def number_of_edges(self):
        return sum(len(n) for n in self._succ.values())


# --------------------------------------------------------------------------- a3
def build_deBruijn_graph(
    pair_frequency,
    normalise=False,
    remove_N=False,
    edge_weight_threshold=None,
    keep_top_k=None,
    create_all_kmers=True,
    disable_tqdm=True,
):
    """src/utils.py:87-165.  Mutates pair_frequency in place when normalising,
    as the reference does (:116-117)."""
    if len(pair_frequency) == 0:
        raise IndexError("list index out of range")  # utils.py:122 on an empty dict
    if normalise:
        out_count = defaultdict(int)
        for (u, _), c in pair_frequency.items():
            out_count[u] += c
        for pair in pair_frequency:
            pair_frequency[pair] /= out_count[pair[0]]

    g = OrderedDiGraph()
    k = len(next(iter(pair_frequency))[0])
    pad = "N" * k
    if create_all_kmers:
        for t in product(ALPHABET, repeat=k):
            g.add_node("".join(t))
    g.add_node(pad)
    for (u, v), w in pair_frequency.items():
        if edge_weight_threshold is None or w > edge_weight_threshold:
            g.add_edge(u, v, w)

    if keep_top_k is not None:
        weights = [w for _, _, w in g.edges(data=True)]
        cut = np.quantile(weights, 1 - keep_top_k)
        for u, v in [(u, v) for u, v, w in g.edges(data=True) if w < cut]:
            g.remove_edge(u, v)
    if remove_N:
        g.remove_node(pad)
    return g


# --------------------------------------------------------------------------- a4
def from_networkx(g):
    """torch_geometric.utils.from_networkx (PyG 2.3.1) as used at
    gnn_tasks/utils.py:37 and gnn_tasks_miniBatch/dataloader.py:78 -- PARITY
    UNPINNED (third party, no reference test).  Nodes -> range(N) in g.nodes()
    order; edges in g.edges() order; python-float weights -> float32."""
    nodes = g.nodes()
    index = {n: i for i, n in enumerate(nodes)}
    src, dst, w = [], [], []
    for u, v, wt in g.edges(data=True):
        src.append(index[u])
        dst.append(index[v])
        w.append(wt)
    edge_index = np.array([src, dst], dtype=np.int64).reshape(2, -1)
    return {
        "edge_index": edge_index,
        "weight": np.asarray(w, dtype=np.float64).astype(np.float32),
        "num_nodes": len(nodes),
        "nodes": nodes,
    }


# --------------------------------------------------------------------------- a5
def node_embedding_initial(graph, method="onehot", k=3, small_k=None, random_size=4):
    """gnn_utils.py:428-493.  kmer_frequency: count the k-s+1 overlapping
    s-mers of every node label (:462-471), drop never-seen columns (:474-481),
    scale by 1/(k-s+1) (:484-485, scipy multiplies by the reciprocal), dense
    float64 (:488)."""
    labels = graph.nodes() if hasattr(graph, "nodes") else list(graph)
    labels = list(labels)
    if method == "onehot":
        eye = np.eye(len(labels))
        return [eye[i] for i in range(len(labels))]
    if method == "random":
        return np.random.rand(len(labels), random_size)
    if method != "kmer_frequency":
        raise ValueError("Invalid method for node embedding initialisation")
    if small_k is None:
        raise ValueError("small_k must be provided for kmer_frequency method")
    col = {"".join(t): i for i, t in enumerate(product(ALPHABET, repeat=small_k))}
    counts = np.zeros((len(labels), len(col)), dtype=np.int64)
    for i, kmer in enumerate(labels):
        for j in range(len(kmer) - small_k + 1):
            counts[i, col[kmer[j : j + small_k]]] += 1
    keep = np.nonzero(counts.sum(axis=0) > 0)[0]
    m = len(labels[-1]) - small_k + 1
    return counts[:, keep].astype(np.float64) * (1.0 / m)


# --------------------------------------------------------------------------- a6
def cosine_similarity_to_edge_index_weight(X, chunk_size=1024, threshold=0.0, keep_top_k=None, row_blocks=None):
    """gnn_utils.py:240-322 -- blocked float64 cosine, strict upper triangle,
    strict ``> threshold``, blocks appended in (i, j>=i) order and row-major
    inside a block; optional global top-n by np.argpartition (:311-317).
    ``row_blocks`` (not in the reference) limits the outer loop to the first
    blocks: bench.py times that bounded sample and scales by block pairs."""
    X = np.asarray(X)
    n = X.shape[0]
    nb = (n + chunk_size - 1) // chunk_size
    src, dst, wts = [], [], []
    for bi in range(nb if row_blocks is None else min(nb, row_blocks)):
        a = X[bi * chunk_size : min((bi + 1) * chunk_size, n)]
        na = np.linalg.norm(a, axis=1).reshape(-1, 1)
        for bj in range(bi, nb):
            b = X[bj * chunk_size : min((bj + 1) * chunk_size, n)]
            nbn = np.linalg.norm(b, axis=1).reshape(1, -1)
            with np.errstate(divide="ignore", invalid="ignore"):
                sim = (a @ b.T) / (na * nbn)
            np.nan_to_num(sim, copy=False)
            if bi == bj:
                sim = np.triu(sim, k=1)
            s, t = np.nonzero(sim > threshold)
            src.append(s + bi * chunk_size)
            dst.append(t + bj * chunk_size)
            wts.append(sim[s, t])
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    wts = np.concatenate(wts)
    if keep_top_k is not None:
        top_n = min(int(n**2 * keep_top_k), wts.shape[0])
        idx = np.argpartition(wts, -top_n)[-top_n:]
        src, dst, wts = src[idx], dst[idx], wts[idx]
    return np.stack([src, dst], axis=0).astype(np.int64), wts


# --------------------------------------------------------------------------- a7 / a8 glue
def full_batch_graph(graph, labels, labels_to_predict=None):
    """gnn_tasks/utils.py:8-51 without PyG: Data fields as a dict."""
    d = from_networkx(graph)
    d["x"] = np.asarray(labels, dtype=np.float64).astype(np.float32)
    d["y"] = [np.asarray(l, dtype=np.float64).astype(np.float32) for l in (labels_to_predict or [])]
    return d


def build_reference_graph(sequences, k, small_k="2", stride=1, create_all_kmers=False,
                          edges_threshold=0.0, edges_keep_top_k=None):
    """kmer_ssgnn.py:162-213 + gnn_tasks/sampling_task.py:128-161: the whole
    hot path end to end, full-batch flavour.  Returns a dict of numpy arrays."""
    _, pf = weighted_directed_edges(sequences, k=k, stride=stride, inlcudeUNK=False)
    g = build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=create_all_kmers)
    kf = [node_embedding_initial(g, "kmer_frequency", k, int(s)) for s in small_k.split(",")]
    out = full_batch_graph(g, kf[0], kf)
    for i, x in enumerate(kf):
        ei, w = cosine_similarity_to_edge_index_weight(np.vstack(x), threshold=edges_threshold,
                                                       keep_top_k=edges_keep_top_k)
        out[f"edge_index_KF{i}"] = ei
        out[f"edge_weight_KF{i}"] = w
    out["kf_labels"] = kf
    return out
