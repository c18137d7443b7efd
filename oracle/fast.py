"""Vectorised CPU restatement of the hot path (test infrastructure only).

Same results as ``oracle/faithful.py`` (and therefore as the reference, see
tests/test_oracle_*.py) but in numpy array form so that it finishes in seconds
at sizes where the reference's pure-Python loops need minutes.  It also states
the *exact-integer* form of the KF cosine test that the CUDA path implements
(SURVEY.md section 0, fact 3): features are small integer counts c/m, so
cos = dot / sqrt(n_i n_j) with integer dot, n_i, n_j.

Reference lines followed: src/utils.py:15-31,52-84,87-165;
src/representations/gnn_common/gnn_utils.py:240-322,428-493.
"""
from fractions import Fraction
from math import isqrt

import numpy as np

SEP = 10  # '\n' between reads in the byte stream
_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _LUT[_c] = _i
_LUT[ord("N")] = 4
_LUT[SEP] = 5


def synthetic_reads(n_reads, length=150, seed=0):
    """i.i.d. uniform bases, ``default_rng(seed).integers(0,4,(R,L))``
    mapped through b"ACGT".  Returns a uint8 [R, L] ASCII array."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=(n_reads, length), dtype=np.uint8)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


def reads_to_strings(arr):
    return [bytes(r).decode("ascii") for r in np.asarray(arr, dtype=np.uint8)]


def to_stream(sequences):
    """ASCII byte stream with one SEP after every read."""
    if isinstance(sequences, np.ndarray) and sequences.dtype == np.uint8 and sequences.ndim == 2:
        r, l = sequences.shape
        out = np.empty((r, l + 1), dtype=np.uint8)
        out[:, :l] = sequences
        out[:, l] = SEP
        return out.reshape(-1)
    joined = "\n".join(sequences) + "\n" if len(sequences) else ""
    return np.frombuffer(joined.encode("ascii"), dtype=np.uint8)


def decode_kmers(codes, k):
    codes = np.asarray(codes, dtype=np.int64)
    out = np.empty((codes.shape[0], k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = np.frombuffer(b"ACGT", dtype=np.uint8)[(codes >> (2 * (k - 1 - j))) & 3]
    return [bytes(r).decode("ascii") for r in out]


# --------------------------------------------------------------------------- a1 + a2
def count_pairs(sequences, k, stride=1):
    """weighted_directed_edges(..., inlcudeUNK=False) (src/utils.py:52-84) in
    array form.  Returns (u, v, count, first_pos) with one row per distinct
    (u, v) pair of consecutive surviving k-mers, rows in dict insertion order
    (= ascending first_pos, the stream offset of u's first base)."""
    stream = to_stream(sequences)
    codes = _LUT[stream]
    if (codes == 255).any():
        raise ValueError("byte outside ACGTN in input")
    n = codes.shape[0]
    empty = (np.zeros(0, np.int64),) * 4
    if n < k:
        return empty
    bad = np.concatenate([[0], np.cumsum(codes > 3)])
    sep = np.concatenate([[0], np.cumsum(codes == 5)])
    p = np.arange(n - k + 1)
    clean = (bad[p + k] - bad[p]) == 0          # k bases, no N, no separator
    if stride != 1:
        is_sep = np.nonzero(codes == 5)[0]
        read_start = np.zeros(n, dtype=np.int64)
        read_start[is_sep[:-1] + 1] = is_sep[:-1] + 1
        read_start = np.maximum.accumulate(read_start)
        clean &= ((p - read_start[p]) % stride) == 0
    vp = np.nonzero(clean)[0]
    if vp.shape[0] < 2:
        return empty
    code = np.zeros(n - k + 1, dtype=np.int64)
    for j in range(k):
        code += codes[j : n - k + 1 + j].astype(np.int64) << (2 * (k - 1 - j))
    rid = sep[vp]
    same = rid[1:] == rid[:-1]
    a, b = vp[:-1][same], vp[1:][same]
    u, v = code[a], code[b]
    key = (u << (2 * k)) | v
    _, first_idx, cnt = np.unique(key, return_index=True, return_counts=True)
    order = np.argsort(first_idx, kind="stable")
    first_idx, cnt = first_idx[order], cnt[order]
    return u[first_idx], v[first_idx], cnt.astype(np.int64), a[first_idx].astype(np.int64)


# --------------------------------------------------------------------------- a3 + a4
def build_db(u, v, cnt, k, normalise=True, create_all_kmers=False,
             edge_weight_threshold=None, keep_top_k=None):
    """build_deBruijn_graph + from_networkx (src/utils.py:87-165, SURVEY.md
    Appendix B) on arrays in dict order.  Returns (node_codes int64[N],
    edge_index int64[2,E], weight float64[E])."""
    u = np.asarray(u, np.int64)
    v = np.asarray(v, np.int64)
    if u.shape[0] == 0:
        raise IndexError("list index out of range")
    w = np.asarray(cnt, np.float64)
    if normalise:
        uu, inv = np.unique(u, return_inverse=True)
        tot = np.bincount(inv, weights=np.asarray(cnt, np.float64))
        w = w / tot[inv]
    if edge_weight_threshold is not None:
        keep = w > edge_weight_threshold
        u, v, w = u[keep], v[keep], w[keep]
    if create_all_kmers:
        nodes = np.arange(4**k, dtype=np.int64)
    else:
        seq = np.stack([u, v], axis=1).reshape(-1)
        uniq, first = np.unique(seq, return_index=True)
        nodes = uniq[np.argsort(first, kind="stable")]
    rank = np.full(4**k, -1, dtype=np.int64)
    rank[nodes] = np.arange(nodes.shape[0])
    order = np.argsort(rank[u], kind="stable")
    u, v, w = u[order], v[order], w[order]
    if keep_top_k is not None and w.shape[0]:
        cut = np.quantile(w, 1 - keep_top_k)
        keep = ~(w < cut)
        u, v, w = u[keep], v[keep], w[keep]
    return nodes, np.stack([rank[u], rank[v]], axis=0), w


# --------------------------------------------------------------------------- a5
def subkmer_counts(node_codes, k, s):
    """Integer sub-k-mer counts [N, 4^s] (gnn_utils.py:462-471); column index is
    the base-4 value of the s-mer with A,C,G,T = 0..3 (:458-459)."""
    node_codes = np.asarray(node_codes, np.int64)
    n = node_codes.shape[0]
    out = np.zeros((n, 4**s), dtype=np.uint8)
    rows = np.arange(n)
    for j in range(k - s + 1):
        col = (node_codes >> (2 * (k - s - j))) & (4**s - 1)
        np.add.at(out, (rows, col), 1)
    return out


def features_from_counts(counts, k, s):
    """Column drop + scaling of gnn_utils.py:474-488: float64 [N, D']."""
    keep = np.nonzero(counts.sum(axis=0, dtype=np.int64) > 0)[0]
    return counts[:, keep].astype(np.float64) * (1.0 / (k - s + 1)), keep


# --------------------------------------------------------------------------- a6 (exact integer form)
def min_dot_table(threshold, p_max):
    """dmin[P] = smallest integer d with d / sqrt(P) > threshold, exactly
    (threshold is a float64, i.e. a dyadic rational).  dmin[0] is unreachable."""
    if threshold < 0:
        raise ValueError("negative KF threshold is not supported")
    fr = Fraction(threshold)
    a2, b2 = fr.numerator**2, fr.denominator**2
    tab = np.empty(p_max + 1, dtype=np.int64)
    for p in range(p_max + 1):
        tab[p] = isqrt((a2 * p) // b2) + 1
    tab[0] = np.iinfo(np.int32).max
    return tab


def kf_candidates(counts, threshold=0.0, chunk_size=1024):
    """All pairs s<t whose exact cosine exceeds ``threshold``, in the
    reference's emission order (block (i, j>=i), row-major inside a block,
    gnn_utils.py:262-301).  Returns (src, dst, dot, P=n_i*n_j) int64 arrays."""
    c = np.ascontiguousarray(counts).astype(np.float32)
    n = c.shape[0]
    sq = (counts.astype(np.int64) ** 2).sum(axis=1)
    dmin = min_dot_table(threshold, int(sq.max()) ** 2 if n else 0)
    nb = (n + chunk_size - 1) // chunk_size
    S, T, Dt, P = [], [], [], []
    for bi in range(nb):
        a0, a1 = bi * chunk_size, min((bi + 1) * chunk_size, n)
        for bj in range(bi, nb):
            b0, b1 = bj * chunk_size, min((bj + 1) * chunk_size, n)
            dot = np.rint(c[a0:a1] @ c[b0:b1].T).astype(np.int64)
            prod = sq[a0:a1, None] * sq[None, b0:b1]
            ok = dot >= dmin[prod]
            if bi == bj:
                ok = np.triu(ok, k=1)
            s, t = np.nonzero(ok)
            S.append(s + a0)
            T.append(t + b0)
            Dt.append(dot[s, t])
            P.append(prod[s, t])
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return cat(S), cat(T), cat(Dt), cat(P)


def kf_weight(dot, prod):
    """float64 cosine from the exact integers -- what the CUDA path emits."""
    return np.asarray(dot, np.float64) / np.sqrt(np.asarray(prod, np.float64))


def top_n_budget(n_nodes, keep_top_k, n_candidates):
    """gnn_utils.py:311-312 including the quirk that 0 keeps everything
    (``[-0:]`` is the whole array)."""
    if keep_top_k is None:
        return n_candidates
    top_n = min(int(n_nodes**2 * keep_top_k), n_candidates)
    return n_candidates if top_n == 0 else top_n


def tie_ratio(tie_budget, tie_total):
    """r = ceil(B 2^64 / T): the fixed-point slope of the spread rule below (0: nothing to cut)."""
    if not (0 < tie_budget < tie_total):
        return 0
    return -((-tie_budget << 64) // tie_total)


def spread_keep(tie_total, tie_budget):
    """Which of the T members of the cut-off class, numbered g = 0..T-1 in emission order, the CUDA path keeps when B
    of them fit the budget: member g iff floor((g+1) r / 2^64) > floor(g r / 2^64), r = ceil(B 2^64 / T).  Exactly B
    pass, evenly spread over the emission order -- so that every row-block shard of a multi-GPU build keeps its share
    and the choice does not depend on how the order is cut into shards or chunks.  (The reference keeps whichever
    members np.argpartition's introselect happens to leave in the top-n slots, gnn_utils.py:311-317: any B of the T
    are an equally faithful answer, see tie_aware_compare.)  Returns a bool mask of length T."""
    r = tie_ratio(tie_budget, tie_total)
    steps = (np.arange(tie_total + 1, dtype=object) * r) >> 64
    return np.asarray(steps[1:] > steps[:-1], dtype=bool)


def kf_edges_exact(counts, threshold=0.0, keep_top_k=None, chunk_size=1024):
    """The CUDA path's KF contract: candidates by the exact test, then, when the
    global top-n budget binds, every pair strictly above the cut-off score class
    plus the members of the cut-off class picked by ``spread_keep``.  Output keeps
    emission order.  Returns (edge_index int64[2,E], weight float64[E])."""
    s, t, dot, prod = kf_candidates(counts, threshold, chunk_size)
    n = counts.shape[0]
    budget = top_n_budget(n, keep_top_k, s.shape[0])
    if budget < s.shape[0]:
        # exact ordering of score classes: compare dot^2 / P as rationals
        cls, inv = np.unique(np.stack([dot, prod], 1), axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        val = [Fraction(int(d) ** 2, int(p)) for d, p in cls]
        order = sorted(range(len(val)), key=lambda i: val[i], reverse=True)
        cnt = np.bincount(inv, minlength=len(val))
        kept_cls = np.zeros(len(val), dtype=bool)
        taken, i = 0, 0
        cut_val, room = None, 0
        while i < len(order):
            j = i
            grp = []
            while j < len(order) and val[order[j]] == val[order[i]]:
                grp.append(order[j])
                j += 1
            size = int(cnt[grp].sum())
            if taken + size <= budget:
                kept_cls[grp] = True
                taken += size
                i = j
                if taken == budget:
                    break
            else:
                cut_val, room = grp, budget - taken
                break
        keep = kept_cls[inv]
        if cut_val is not None and room > 0:
            tie = np.nonzero(np.isin(inv, cut_val))[0]
            keep[tie[spread_keep(tie.shape[0], room)]] = True
        s, t, dot, prod = s[keep], t[keep], dot[keep], prod[keep]
    return np.stack([s, t], axis=0).astype(np.int64), kf_weight(dot, prod)


# --------------------------------------------------------------------------- whole path
def build_graph(sequences, k, small_k="2", stride=1, create_all_kmers=False,
                edges_threshold=0.0, edges_keep_top_k=None):
    """kmer_ssgnn.py:162-213 + sampling_task.py:128-161 in array form."""
    u, v, cnt, _ = count_pairs(sequences, k, stride)
    nodes, ei, w = build_db(u, v, cnt, k, normalise=True, create_all_kmers=create_all_kmers)
    out = {"node_codes": nodes, "edge_index": ei, "weight": w.astype(np.float32), "num_nodes": len(nodes)}
    for i, s in enumerate(int(x) for x in small_k.split(",")):
        c = subkmer_counts(nodes, k, s)
        x, keep = features_from_counts(c, k, s)
        out[f"counts_{i}"] = c
        out[f"y_{i}"] = x.astype(np.float32)
        kei, kw = kf_edges_exact(c[:, keep], edges_threshold, edges_keep_top_k)
        out[f"edge_index_KF{i}"] = kei
        out[f"edge_weight_KF{i}"] = kw
    out["x"] = out["y_0"]
    return out


# --------------------------------------------------------------------------- comparison helpers
def tie_aware_compare(ei_a, w_a, ei_b, w_b, rel=1e-9):
    """SURVEY.md 8(c) parity metric for a top-n-bound edge set: a = ours,
    b = reference.  Returns a dict of diagnostics; ``ok`` is the verdict."""
    ka = ei_a[0].astype(np.int64) << 32 | ei_a[1].astype(np.int64)
    kb = ei_b[0].astype(np.int64) << 32 | ei_b[1].astype(np.int64)
    sa, sb = set(ka.tolist()), set(kb.tolist())
    c = float(w_b.min()) if w_b.shape[0] else 0.0
    above_a = set(ka[w_a > c * (1 + rel)].tolist())
    above_b = set(kb[w_b > c * (1 + rel)].tolist())
    rest_a = w_a[~(w_a > c * (1 + rel))]
    inter = len(sa & sb)
    union = len(sa | sb)
    res = {
        "E_a": len(sa), "E_b": len(sb), "cutoff": c,
        "jaccard_raw": inter / union if union else 1.0,
        "above_equal": above_a == above_b,
        "rest_in_tie_class": bool(np.all(np.abs(rest_a - c) <= rel * max(c, 1e-300))),
        "jaccard_tie_excluded": (len(above_a & above_b) / len(above_a | above_b)) if (above_a | above_b) else 1.0,
    }
    res["ok"] = res["E_a"] == res["E_b"] and res["above_equal"] and res["rest_in_tie_class"]
    return res
