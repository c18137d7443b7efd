"""Row f4b: per-row kNN KF edges (the faiss_ann mode) against the exact-search restatement oracle/knn.py.

The reference runs the search inside FAISS (absent here; approximate IVFFlat by default), so what is pinned is the
exact search it approximates plus the reference's own post-processing (gnn_utils.py:205-231): identical edge_index
(deterministic tie rule: ascending column), weights within 1e-6 (float32 arithmetic in a different order).
"""
import numpy as np
import pytest
import torch

from oracle import fast, knn

pytestmark = pytest.mark.gpu


def _built(k, s, n_reads, seed=3):
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=seed)
    built = core.build_graph(reads, k, str(s), with_kf=False)
    return built, built.kf_counts[0]


@pytest.mark.parametrize("k,s,n_reads,frac", [
    (4, 2, 400, 0.05),     # D = 16, N = 256: two row blocks, heavy ties
    (6, 2, 4000, 0.01),    # D = 16, N = 4096
    (6, 3, 4000, 0.01),    # D = 64
    (6, 5, 4000, 0.01),    # D = 1024: eight staged chunks per column tile
    (5, 1, 2000, 0.02),    # D = 4 (word loads)
    (7, 5, 30, 0.004),     # ragged N (not every 7-mer occurs): partial row block and column tile
    (6, 4, 4000, 0.3),     # large k: cut classes deep in the ranking; IP rows run out of positive neighbours
])
@pytest.mark.parametrize("distance", ["IP", "L2"])
def test_knn_matches_exact_search(k, s, n_reads, frac, distance):
    from genomic_gnn_b200 import core

    built, kc = _built(k, s, n_reads)
    n = built.db.num_nodes
    k_nn = max(1, int(frac * n))
    got = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance)
    want_ei, want_w = knn.knn_edges(kc.counts.cpu().numpy(), kc.m, k_nn, distance)
    ei = got.edge_index.cpu().numpy()
    assert ei.shape == want_ei.shape, (ei.shape, want_ei.shape)
    assert np.array_equal(ei, want_ei)
    np.testing.assert_allclose(got.weight.cpu().numpy(), want_w, rtol=0, atol=1e-6)
    assert np.all(ei[0] != ei[1]) and np.all(np.diff(ei[0]) >= 0)
    assert np.bincount(ei[0], minlength=n).max() <= k_nn


@pytest.mark.parametrize("k,s,n_reads,frac", [
    (6, 5, 4000, 0.01),      # D = 1024, N = 4096: 16 strip pairs, 16 column tiles
    (7, 5, 30, 0.004),       # ragged N: partial strip pair, partial column tile
    (6, 4, 4000, 0.3),       # D = 256, large k: cuts deep in the ranking, IP rows run out of positive neighbours
    (8, 5, 60000, 0.01),     # BASELINE config 4 size: N = 65,536, D = 1024, k = 655
    (9, 4, 400000, 0.0005),  # N ~ 262k > 65,536: 32-bit shared-memory histograms
])
@pytest.mark.parametrize("distance", ["IP", "L2"])
def test_knn_tcgen05_matches_mma(k, s, n_reads, frac, distance):
    """The inverted-index join and both Gram passes on the tcgen05 CTA-pair kernel (rank histograms with the zero-dot classes by subtraction, pass 2
    by one epilogue group in column order) against the mma.sync kernel: identical edges, order and weights -- whole and
    as a 128-aligned row shard."""
    from genomic_gnn_b200 import core

    built, kc = _built(k, s, n_reads, seed=0 if k >= 8 else 3)
    n = built.db.num_nodes
    k_nn = max(1, int(frac * n))
    a = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, impl="mma")
    b = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, impl="tcgen05")
    assert a.edge_index.shape[1] > 0
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight, b.weight)
    if n <= 200_000:   # the inverted-index join (the default for these shapes)
        c = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, impl="sparse")
        assert torch.equal(a.edge_index, c.edge_index) and torch.equal(a.weight, c.weight)
    if n >= 1024 and k <= 8:
        lo, hi = 384, min(n, 384 + 640)   # a shard that starts inside a strip pair
        fix = (lambda q: q) if distance == "L2" else None
        a1 = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, rows=(lo, hi), qmax_reduce=fix, impl="mma")
        b1 = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, rows=(lo, hi), qmax_reduce=fix, impl="tcgen05")
        assert torch.equal(a1.edge_index, b1.edge_index) and torch.equal(a1.weight, b1.weight)
        c1 = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, rows=(lo, hi), qmax_reduce=fix, impl="sparse")
        assert torch.equal(a1.edge_index, c1.edge_index) and torch.equal(a1.weight, c1.weight)


@pytest.mark.parametrize("distance", ["IP", "L2"])
def test_knn_sparse_on_64_columns(distance):
    """Rows of 64 counts (small_k = 3): too narrow for the tensor path, still sparse enough for the join."""
    from genomic_gnn_b200 import core

    built, kc = _built(6, 3, 4000)
    a = core.kf_knn_edges(kc.counts, kc.sqnorm, 60, distance, impl="mma")
    c = core.kf_knn_edges(kc.counts, kc.sqnorm, 60, distance, impl="sparse")
    assert a.edge_index.shape[1] > 0
    assert torch.equal(a.edge_index, c.edge_index) and torch.equal(a.weight, c.weight)


@pytest.mark.parametrize("hist_mode", [0, 1, 2])
@pytest.mark.parametrize("distance", ["IP", "L2"])
def test_knn_histogram_modes_agree(hist_mode, distance):
    """Pass 1 accumulates in global memory (atomics) or in per-CTA shared memory with 16- or 32-bit bins; the mode is
    chosen by size, so the small cases force each one."""
    from genomic_gnn_b200 import core

    built, kc = _built(6, 2, 4000)
    k_nn = 40
    want = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance)
    got = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, hist_mode=hist_mode)
    assert torch.equal(got.edge_index, want.edge_index) and torch.equal(got.weight, want.weight)


@pytest.mark.parametrize("distance", ["IP", "L2"])
def test_knn_row_shards_concatenate(distance):
    """Rows are independent: 128-aligned row shards (the multi-GPU unit), computed separately and agreeing only on the
    "L2" normaliser, concatenate to the unsharded result."""
    from genomic_gnn_b200 import core

    built, kc = _built(6, 3, 4000)
    n = built.db.num_nodes
    k_nn = 50
    want = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance)
    world = 3
    shards = [core.knn_row_shard(n, world, r) for r in range(world)]
    assert shards[0][0] == 0 and shards[-1][1] == n and all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    qmaxes = []
    if distance == "L2":   # what the ranks exchange: one integer each
        for lo, hi in shards:
            core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, rows=(lo, hi),
                              qmax_reduce=lambda q: qmaxes.append(q) or max(q, 1))
    parts = [core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, rows=(lo, hi),
                               qmax_reduce=(lambda q: max(qmaxes)) if distance == "L2" else None)
             for lo, hi in shards]
    ei = torch.cat([p.edge_index for p in parts], dim=1)
    w = torch.cat([p.weight for p in parts])
    assert torch.equal(ei, want.edge_index) and torch.equal(w, want.weight)


def test_knn_api_and_builder_flag():
    """kf_faiss_to_edge_index_weight on the float matrix the reference passes (np.vstack(labels)), on the labels object,
    and through the mini-batch builder's faiss_ann flag (dataloader.py:108-121: k = int(edges_keep_top_k * N))."""
    import genomic_gnn_b200 as gg
    from genomic_gnn_b200 import builders

    reads = fast.reads_to_strings(fast.synthetic_reads(3000, 150, seed=5))
    _, pf = gg.weighted_directed_edges(reads, k=5, stride=1, inlcudeUNK=False)
    graph = gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    labels = gg.node_embedding_initial(graph, method="kmer_frequency", k=5, small_k=2)
    n = graph.number_of_nodes()
    k_nn = int(0.02 * n)
    ei_a, w_a = gg.kf_faiss_to_edge_index_weight(np.vstack(labels), k=k_nn, index_type="IVFFlat", distance="L2")
    ei_b, w_b = gg.kf_faiss_to_edge_index_weight(labels, k=k_nn, distance="L2")
    assert ei_a.dtype == np.int64 and w_a.dtype == np.float32
    assert np.array_equal(ei_a, ei_b) and np.array_equal(w_a, w_b)
    want_ei, want_w = knn.knn_edges(labels.kc.counts.cpu().numpy(), labels.kc.m, k_nn, "L2")
    assert np.array_equal(ei_a, want_ei)
    np.testing.assert_allclose(w_a, want_w, rtol=0, atol=1e-6)
    _, data = builders.networkx_to_dataloader_mini_batch(
        graph, labels, labels_to_predict=[labels], batch_size=64, kmer_frequency_loss_weight=1.0,
        kmer_freq_labels=[labels], k=5, edges_keep_top_k=0.02, layers_config=["32_KF0"], ss_task="CL", faiss_ann=True,
        faiss_distance="L2")
    rel = data["node", "KF0", "node"]
    assert np.array_equal(rel.edge_index.cpu().numpy(), want_ei)
    assert rel.edge_attr.dtype == torch.float32


def test_knn_full_size_properties_k8():
    """BASELINE config 4 size (N = 65,536, k = 655 per row): size-independent properties + a sampled exact check."""
    from genomic_gnn_b200 import core

    built, kc = _built(8, 2, 60000, seed=0)
    n = built.db.num_nodes
    k_nn = int(0.01 * n)
    counts = kc.counts.cpu().numpy().astype(np.int64)
    sq = (counts * counts).sum(axis=1)
    for distance in ("IP", "L2"):
        got = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance)
        ei, w = got.edge_index.cpu().numpy(), got.weight.cpu().numpy()
        per_row = np.bincount(ei[0], minlength=n)
        assert per_row.max() <= k_nn and np.all(np.diff(ei[0]) >= 0) and np.all(ei[0] != ei[1])
        assert w.min() > 1e-4 and w.max() <= 1.0
        starts = np.concatenate([[0], np.cumsum(per_row)])
        rng = np.random.default_rng(1)
        for i in rng.integers(0, n, 24):
            seg = slice(starts[i], starts[i + 1])
            assert np.all(np.diff(w[seg]) <= 0)            # most similar first
            dots = counts @ counts[i]
            if distance == "IP":
                key = np.where(dots > 0, dots * dots / sq.astype(np.float64), -1.0)
                key[i] = -1.0
                order = np.lexsort((np.arange(n), -key))[:k_nn]
                order = order[key[order] > 0]
            else:
                d = sq + sq[i] - 2 * dots
                key = d.astype(np.float64)
                key[i] = np.inf
                order = np.lexsort((np.arange(n), key))[:k_nn]
                order = order[: per_row[i]]                # the suffix at the global maximum distance is dropped
            assert np.array_equal(ei[1][seg], order), (distance, int(i))
