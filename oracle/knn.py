"""CPU restatement of the reference's per-row kNN KF edges (TEST INFRASTRUCTURE ONLY -- never imported by the product
path; see oracle/__init__.py).

Reference: src/representations/gnn_common/gnn_utils.py:122-237 (kf_faiss_to_edge_index_weight) with its call site
gnn_tasks_miniBatch/dataloader.py:108-121 (k = int(edges_keep_top_k * N)).  The search itself happens inside FAISS
(faiss-gpu 1.7.x in the reference's environment.yml; NOT installed in this image and not under /root/reference), by
default through an approximate IVFFlat index trained with k-means.  What is restated here is the exact search that
index approximates (faiss.IndexFlatL2 / IndexFlatIP semantics: squared Euclidean distances / inner products of the
L2-normalised rows, best first) followed by the reference's own post-processing, line by line.  PARITY UNPINNED: no
FAISS here, no golden vectors in the reference; the anchor is the published definition of the flat indexes, and
tests/test_oracle.py cross-checks this file against scikit-learn's brute-force NearestNeighbors (an independent
implementation of the same exact search).

Two deliberate, documented choices where the reference's behaviour is an artefact of FAISS:
  * the reference drops column 0 of the (k + 1)-wide result assuming it is the query itself (gnn_utils.py:207-211);
    with duplicate rows FAISS may return a duplicate there and keep the self match.  Here the self match is excluded
    explicitly.
  * ties are ordered by ascending column (FAISS's order inside a tie is heap-dependent).
"""
import numpy as np


def knn_edges(counts, m, k, distance="L2"):
    """counts: integer [N, D] sub-k-mer counts of the nodes (X = counts * (1 / m) is what the reference passes).
    Returns (edge_index int64 [2, E], edge_weight float32 [E]) like gnn_utils.py:233-237."""
    c = np.asarray(counts).astype(np.int64)
    n = c.shape[0]
    sq = (c * c).sum(axis=1)
    cf = c.astype(np.float64)
    G = np.rint(cf @ cf.T).astype(np.int64)                         # BLAS; exact: small integers
    cols = np.arange(n)
    if distance == "IP":
        # faiss.normalize_L2(X) then inner product (gnn_utils.py:170-172): the cosine.  Order by dot^2 / sq_j (exact in
        # float64: distinct rationals of these sizes never round to the same double).
        with np.errstate(divide="ignore", invalid="ignore"):
            key = np.where(sq[None, :] > 0, (G * G) / sq[None, :].astype(np.float64), 0.0)
        key[cols, cols] = -1.0
        order = np.argsort(-key, axis=1, kind="stable")[:, :k]    # stable: ties by ascending column
        with np.errstate(divide="ignore", invalid="ignore"):
            cos = G / np.sqrt((sq[:, None] * sq[None, :]).astype(np.float64))
        D = np.take_along_axis(np.nan_to_num(cos), order, axis=1).astype(np.float32)
        valid = np.take_along_axis(key, order, axis=1) > 0          # a zero dot is similarity 0: filtered below anyway
        D = np.where(valid, D, np.float32(0.0))
    elif distance == "L2":
        dist = sq[:, None] + sq[None, :] - 2 * G                    # m^2 * squared Euclidean distance of the rows of X
        key = dist.astype(np.float64)
        key[cols, cols] = np.inf
        order = np.argsort(key, axis=1, kind="stable")[:, :k]
        D = (np.take_along_axis(dist, order, axis=1) / float(m * m)).astype(np.float32)
        max_distance = np.max(D)                                    # gnn_utils.py:214-216
        D = 1 - (D / max_distance)
    else:
        raise ValueError(distance)
    I = order
    D = D.clip(max=1.0)                                             # :217
    nonzero_mask = D.flatten() > 0.0001                             # :221
    source = np.repeat(np.arange(n), k)[nonzero_mask]               # :222
    target = I.flatten()[nonzero_mask]
    edge_weight = D.flatten()[nonzero_mask]
    return np.stack([source, target], axis=0).astype(np.int64), edge_weight.astype(np.float32)
