"""Pin the CPU oracle against the reference's own outputs (tests/golden made by
oracle/make_golden.py from /root/reference) -- CPU only."""
import os

import numpy as np
import pytest

from oracle import faithful, fast, ref_import
from tests._util import (GOLDEN, assert_kf_threshold_parity, edge_keys, load_golden, small_goldens)


@pytest.fixture(params=small_goldens(), ids=lambda p: p.split("small_")[-1][:-4])
def golden(request):
    return load_golden(request.param)


def _budget_binds(n_nodes, topk, n_ref_edges):
    if topk is None:
        return False
    n = int(n_nodes**2 * topk)
    return n != 0 and n_ref_edges == n


def test_golden_files_present():
    assert len(small_goldens()) >= 5


def test_faithful_matches_reference(golden):
    g = golden
    _, pf = faithful.weighted_directed_edges(g["reads"], k=g["k"], stride=g["stride"], inlcudeUNK=False)
    assert [p[0] for p in pf] == g["pair_u"].tolist()
    assert [p[1] for p in pf] == g["pair_v"].tolist()
    assert list(pf.values()) == g["pair_count"].tolist()
    graph = faithful.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=g["create_all_kmers"])
    d = faithful.from_networkx(graph)
    assert d["nodes"] == g["nodes"].tolist()
    assert np.array_equal(d["edge_index"], g["edge_index"])
    assert np.array_equal(d["weight"], g["weight64"].astype(np.float32))
    for s in g["small_ks"]:
        x = faithful.node_embedding_initial(graph, "kmer_frequency", g["k"], s)
        assert x.dtype == np.float64 and np.array_equal(x, g[f"x_s{s}"])
        for j, (thr, topk) in enumerate(g["kf_settings"]):
            ei, w = faithful.cosine_similarity_to_edge_index_weight(np.vstack(x), threshold=thr, keep_top_k=topk)
            assert np.array_equal(ei, g[f"kf_s{s}_{j}_edge_index"])
            assert np.array_equal(w, g[f"kf_s{s}_{j}_weight"])


def test_fast_matches_reference(golden):
    g = golden
    k = g["k"]
    u, v, cnt, first = fast.count_pairs(g["reads"], k, g["stride"])
    assert fast.decode_kmers(u, k) == g["pair_u"].tolist()
    assert fast.decode_kmers(v, k) == g["pair_v"].tolist()
    assert cnt.tolist() == g["pair_count"].tolist()
    assert np.all(np.diff(first) > 0)
    nodes, ei, w = fast.build_db(u, v, cnt, k, normalise=True, create_all_kmers=g["create_all_kmers"])
    assert fast.decode_kmers(nodes, k) == g["nodes"].tolist()
    assert np.array_equal(ei, g["edge_index"])
    assert np.array_equal(w, g["weight64"])
    for s in g["small_ks"]:
        counts = fast.subkmer_counts(nodes, k, s)
        x, keep = fast.features_from_counts(counts, k, s)
        assert np.array_equal(x, g[f"x_s{s}"])
        for j, (thr, topk) in enumerate(g["kf_settings"]):
            ei_ref, w_ref = g[f"kf_s{s}_{j}_edge_index"], g[f"kf_s{s}_{j}_weight"]
            ei_o, w_o = fast.kf_edges_exact(counts[:, keep], thr, topk)
            assert np.all(ei_o[0] < ei_o[1])
            if _budget_binds(len(nodes), topk, ei_ref.shape[1]):
                r = fast.tie_aware_compare(ei_o, w_o, ei_ref, w_ref)
                assert r["ok"], r
            else:
                assert_kf_threshold_parity(ei_o, w_o, ei_ref, w_ref, thr)
                if topk is None and thr in (0.0, 0.8):
                    # no equality-band artefacts for these thresholds: exact, ordered
                    assert np.array_equal(ei_o, ei_ref)


def test_min_dot_table_is_exact():
    tab = fast.min_dot_table(0.8, 49 * 49)
    # cos == 0.8 exactly (dot=4, P=25) must NOT pass: fl(0.8) > 4/5
    assert tab[25] == 5
    assert tab[100] == 9          # 8/10 == 0.8 rejected, 9/10 passes
    tab0 = fast.min_dot_table(0.0, 16)
    assert np.all(tab0[1:] == 1)
    with pytest.raises(ValueError):
        fast.min_dot_table(-0.1, 4)


def test_top_n_budget_quirks():
    assert fast.top_n_budget(100, None, 77) == 77
    assert fast.top_n_budget(100, 0.01, 1000) == 100
    assert fast.top_n_budget(100, 0.01, 50) == 50
    assert fast.top_n_budget(5, 0.01, 9) == 9      # int(25*0.01)=0 -> [-0:] keeps all


def test_empty_and_short_reads():
    for reads in ([], [""], ["AC"], ["ACG"], ["NNNNNN"], ["ACGNACG"]):
        u, v, c, f = fast.count_pairs(reads, 3, 1)
        _, pf = faithful.weighted_directed_edges(reads, k=3, stride=1, inlcudeUNK=False)
        assert len(pf) == u.shape[0]
    with pytest.raises(IndexError):
        faithful.build_deBruijn_graph({}, normalise=True)
    with pytest.raises(ValueError):
        fast.count_pairs(["ACGX"], 3)


def test_internal_N_bridges_non_overlapping_pair():
    # utils.py:31,79-82: k-mers holding N are deleted before pairing, so the
    # k-mer before an N run is paired with the first clean k-mer after it.
    _, pf = faithful.weighted_directed_edges(["ACGTNTTGA"], k=3, stride=1, inlcudeUNK=False)
    assert ("CGT", "TTG") in pf
    u, v, c, f = fast.count_pairs(["ACGTNTTGA"], 3)
    assert list(zip(fast.decode_kmers(u, 3), fast.decode_kmers(v, 3))) == list(pf.keys())


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_live_reference_random_cross_check():
    ref = ref_import.load()
    rng = np.random.default_rng(7)
    for k, stride in [(3, 1), (4, 1), (4, 2), (5, 1)]:
        reads = []
        for _ in range(120):
            ln = int(rng.integers(0, 50))
            s = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
            s = "".join(("N" if rng.random() < 0.04 else ch) for ch in s) + "N" * int(rng.integers(0, 4))
            reads.append(s)
        with ref_import.quiet():
            _, pf = ref.weighted_directed_edges(reads, k=k, stride=stride, inlcudeUNK=False)
        u, v, cnt, _ = fast.count_pairs(reads, k, stride)
        assert list(zip(fast.decode_kmers(u, k), fast.decode_kmers(v, k))) == list(pf.keys())
        assert cnt.tolist() == list(pf.values())
        for cak in (False, True):
            for thr, topk in ((None, None), (1.0 / 4**k, 0.3)):
                with ref_import.quiet():
                    gr = ref.build_deBruijn_graph(dict(pf), normalise=True, remove_N=True, create_all_kmers=cak,
                                                  edge_weight_threshold=thr, keep_top_k=topk)
                nodes, ei, w = fast.build_db(u, v, cnt, k, True, cak, thr, topk)
                assert fast.decode_kmers(nodes, k) == list(gr.nodes())
                idx = {n: i for i, n in enumerate(gr.nodes())}
                ref_ei = np.array([[idx[a], idx[b]] for a, b in gr.edges()], dtype=np.int64).T.reshape(2, -1)
                assert np.array_equal(ref_ei, ei)
                assert np.array_equal(np.array([d for _, _, d in gr.edges(data="weight")]), w)


def test_random_walk_law_matches_reference_statistics():
    """Row f3 oracle pin: the exact walk / window laws of oracle/rw.py against statistics of the REAL reference sampler
    (gnn_utils.py:514-600) frozen by oracle/make_golden_rw.py -- 32,000 decoded walks with p=0.5, q=2.0 and the
    pairs-per-walk histogram of 20,000 single walks."""
    from oracle import rw

    z = np.load(os.path.join(GOLDEN, "rw_reference.npz"))
    n = int(z["num_nodes"])
    nbrs = rw.successors(z["edge_index"], z["weight"], n)
    counts = {tuple(int(x) for x in k): int(c) for k, c in zip(z["walk_keys"], z["walk_counts"])}
    per_start = {}
    for (a, _, _), c in counts.items():
        per_start[a] = per_start.get(a, 0) + c
    assert len(per_start) == n
    chi2, dof = 0.0, 0
    for a in range(n):
        law = rw.walk_law(nbrs, a, 2, p=float(z["p"]), q=float(z["q"]))
        assert abs(sum(law.values()) - 1.0) < 1e-12
        assert set(k for k in counts if k[0] == a) <= set(law)       # the reference never takes an impossible walk
        for walk, pr in law.items():
            exp = per_start[a] * pr
            got = counts.get(walk, 0)
            assert abs(got - exp) <= 5.0 * np.sqrt(exp * (1 - pr)) + 1.0, (walk, got, exp)
            if exp >= 5:
                chi2 += (got - exp) ** 2 / exp
                dof += 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
    hist = {int(k): int(c) for k, c in z["pairs_per_walk"]}
    total = sum(hist.values())
    law = rw.pairs_per_walk_law(5, 3)
    assert set(hist) <= set(law)
    for k, pr in law.items():
        exp = total * pr
        assert abs(hist.get(k, 0) - exp) <= 5.0 * np.sqrt(exp * (1 - pr)) + 1.0, (k, hist.get(k, 0), exp)
    # the line-by-line restatement follows the same law
    import random

    r = random.Random(3)
    seen = {}
    for _ in range(300):
        _, walks = rw.sample(z["edge_index"], z["weight"], n, n, num_walks=2, walk_length=2, window_size=1,
                             p=float(z["p"]), q=float(z["q"]), rng=r)
        for w in walks:
            seen[tuple(w)] = seen.get(tuple(w), 0) + 1
    for a in range(n):
        law = rw.walk_law(nbrs, a, 2, p=float(z["p"]), q=float(z["q"]))
        for walk, pr in law.items():
            exp = 600 * pr
            assert abs(seen.get(walk, 0) - exp) <= 5.0 * np.sqrt(exp * (1 - pr)) + 1.0


def test_knn_tables_rank_like_the_exact_search():
    """Host tables of the kNN KF path (core.knn_tables) against oracle/knn.py's order on a small graph: sorting every
    row by (rank, column) must reproduce the oracle's neighbour lists for both metrics."""
    from genomic_gnn_b200.core import knn_tables
    from oracle import knn

    k, s = 4, 2
    n, d = 4**k, 4**s
    codes = np.arange(n)
    counts = np.zeros((n, d), dtype=np.int64)
    for j in range(k - s + 1):
        np.add.at(counts, (codes, (codes >> (2 * (k - s - j))) & (d - 1)), 1)
    sq = (counts * counts).sum(axis=1)
    G = counts @ counts.T
    k_nn = 12
    for metric in ("IP", "L2"):
        sqid, tab, rank_q, n_ranks = knn_tables(tuple(int(v) for v in np.unique(sq)), metric)
        ranks = tab[sqid[sq][None, :].repeat(n, 0), G].astype(np.int64)
        ranks[codes, codes] = 0xFFFF
        assert ranks[ranks != 0xFFFF].max() < n_ranks
        order = np.lexsort((np.broadcast_to(codes, (n, n)), ranks), axis=1)[:, :k_nn]
        keep = np.take_along_axis(ranks, order, axis=1) != 0xFFFF
        if metric == "L2":
            dist = sq[:, None] + np.take_along_axis(np.broadcast_to(sq, (n, n)), order, axis=1) - 2 * np.take_along_axis(G, order, axis=1)
            assert np.array_equal(dist, sq[:, None] + rank_q[np.take_along_axis(ranks, order, axis=1)])
            keep &= dist < dist.max()
        want_ei, _ = knn.knn_edges(counts, k - s + 1, k_nn, metric)
        assert np.array_equal(np.repeat(codes, k_nn)[keep.ravel()], want_ei[0])
        assert np.array_equal(order.ravel()[keep.ravel()], want_ei[1])


def test_knn_oracle_agrees_with_an_independent_exact_search():
    """oracle/knn.py restates FAISS's flat (exact) search; FAISS is not in the image, scikit-learn's brute-force
    NearestNeighbors is an independent implementation of the same definition.  Per row: identical distances of the k
    kept neighbours (float32 round-off apart) and identical neighbour sets above the row's last tie class."""
    from sklearn.neighbors import NearestNeighbors

    from oracle import knn

    rng = np.random.default_rng(7)
    k, s, m = 6, 3, 4
    codes = rng.choice(4**k, size=700, replace=False)
    d = 4**s
    counts = np.zeros((codes.size, d), dtype=np.int64)
    for j in range(k - s + 1):
        np.add.at(counts, (np.arange(codes.size), (codes >> (2 * (k - s - j))) & (d - 1)), 1)
    X = counts * (1.0 / m)
    n, k_nn = X.shape[0], 25
    # L2: squared Euclidean distances, 1 - d / max d, neighbours at the maximum dropped
    ei, w = knn.knn_edges(counts, m, k_nn, "L2")
    nn = NearestNeighbors(n_neighbors=n, algorithm="brute", metric="sqeuclidean").fit(X)
    dist, idx = nn.kneighbors(X)
    dmax = 0.0
    per_row = []
    for i in range(n):
        keep = idx[i] != i
        di, ii = dist[i][keep][:k_nn], idx[i][keep][:k_nn]
        per_row.append((di, ii, dist[i][keep]))
        dmax = max(dmax, di.max())
    starts = np.concatenate([[0], np.cumsum(np.bincount(ei[0], minlength=n))])
    for i in range(n):
        di, ii, alld = per_row[i]
        sim = 1.0 - di / dmax
        kept = sim > 1e-4
        seg = slice(starts[i], starts[i + 1])
        assert seg.stop - seg.start == int(kept.sum())
        np.testing.assert_allclose(w[seg], sim[kept], rtol=0, atol=2e-6)
        strict = di[kept] < di[kept].max() - 1e-9        # below the last tie class the choice is forced
        assert set(ii[kept][strict]) <= set(ei[1][seg])
    # IP: cosine, zero similarities dropped
    ei, w = knn.knn_edges(counts, m, k_nn, "IP")
    Xn = X / np.linalg.norm(X, axis=1, keepdims=True)
    S = Xn @ Xn.T
    np.fill_diagonal(S, -1.0)
    starts = np.concatenate([[0], np.cumsum(np.bincount(ei[0], minlength=n))])
    for i in range(n):
        top = np.sort(S[i])[::-1][:k_nn]
        top = top[top > 1e-4]
        seg = slice(starts[i], starts[i + 1])
        np.testing.assert_allclose(w[seg], np.minimum(top, 1.0), rtol=0, atol=2e-6)
        assert np.all(np.abs(S[i][ei[1][seg]] - w[seg]) < 2e-6)   # every kept edge carries its true cosine


def test_knn_rank_tables_order_like_exact_arithmetic():
    """core.knn_tables for arbitrary sets of squared norms: equal rank <=> equal similarity, lower rank <=> more
    similar, in exact rational arithmetic ("IP": dot / sqrt(sq_j); "L2": sq_j - 2 dot), and impossible (dot, sq)
    combinations (Cauchy-Schwarz) / zero dots ("IP") are marked as never-a-neighbour."""
    from fractions import Fraction

    from hypothesis import given, settings, strategies as st

    from genomic_gnn_b200.core import knn_tables

    @settings(max_examples=40, deadline=None)
    @given(st.sets(st.integers(min_value=1, max_value=81), min_size=1, max_size=12))
    def check(sq_set):
        sq_values = tuple(sorted(sq_set))
        for metric in ("IP", "L2"):
            sqid, tab, rank_q, n_ranks = knn_tables(sq_values, metric)
            sq_max = sq_values[-1]
            assert tab.shape == (len(sq_values), sq_max + 1)
            entries = []
            for sq in sq_values:
                for dot in range(sq_max + 1):
                    r = int(tab[sqid[sq], dot])
                    possible = dot * dot <= sq * sq_max and not (metric == "IP" and dot == 0)
                    assert (r != 0xFFFF) == possible
                    if possible:
                        key = Fraction(dot * dot, sq) if metric == "IP" else -(sq - 2 * dot)
                        entries.append((r, key))
                        if metric == "L2":
                            assert rank_q[r] == sq - 2 * dot
            assert {r for r, _ in entries} == set(range(n_ranks)) or not entries
            for (r1, k1) in entries[::7]:
                for (r2, k2) in entries[::5]:
                    assert (r1 < r2) == (k1 > k2) and (r1 == r2) == (k1 == k2)

    check()
