"""CUDA path vs the reference (golden fixtures) and vs the CPU oracle.

All tests call through the reference-signature API / the C ABI of
libgg_b200.so.  Bar: bit-exact counts, node order, DB edge_index and float32
weights; KF edge sets exact outside the |cos - thr| <= 1e-12 equality band,
float32 KF weights bit-exact, float64 within 1e-12 relative.
"""
import hashlib
import glob
import os

import numpy as np
import pytest
import torch

from oracle import fast
from tests._util import GOLDEN, assert_kf_threshold_parity, load_golden, small_goldens, sort_edges

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    import genomic_gnn_b200 as g

    return g


def _sha(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def _budget_binds(n_nodes, topk, n_ref_edges):
    if topk is None:
        return False
    n = int(n_nodes**2 * topk)
    return n != 0 and n_ref_edges == n


# --------------------------------------------------------------------------- reference goldens
@pytest.mark.parametrize("path", small_goldens(), ids=lambda p: os.path.basename(p)[6:-4])
def test_api_matches_reference_golden(gg, path):
    g = load_golden(path)
    k = g["k"]
    item, pf = gg.weighted_directed_edges(g["reads"], k=k, stride=g["stride"], inlcudeUNK=False)
    assert [p[0] for p in pf] == g["pair_u"].tolist()
    assert [p[1] for p in pf] == g["pair_v"].tolist()
    assert list(pf.values()) == g["pair_count"].tolist()
    assert sum(item.values()) == int(g["pair_count"].sum())
    graph = gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=g["create_all_kmers"])
    assert graph.nodes() == g["nodes"].tolist()
    assert graph.number_of_nodes() == len(g["nodes"])
    assert np.array_equal(graph.edge_index.cpu().numpy(), g["edge_index"])
    assert graph.edge_index.dtype == torch.int64 and graph.weight.dtype == torch.float32
    assert np.array_equal(graph.weight.cpu().numpy().view(np.int32), g["weight64"].astype(np.float32).view(np.int32))
    assert np.array_equal(graph.db.weight64.cpu().numpy(), g["weight64"])
    for s in g["small_ks"]:
        lab = gg.node_embedding_initial(graph, method="kmer_frequency", k=k, small_k=s)
        x_ref = g[f"x_s{s}"]
        assert lab.shape == x_ref.shape
        assert np.array_equal(np.asarray(lab), x_ref)
        assert np.array_equal(lab.to_torch_float32().cpu().numpy(), x_ref.astype(np.float32))
        for j, (thr, topk) in enumerate(g["kf_settings"]):
            ei_ref, w_ref = g[f"kf_s{s}_{j}_edge_index"], g[f"kf_s{s}_{j}_weight"]
            for X in (lab, np.vstack(lab)):  # device-resident labels and the plain ndarray the reference passes
                ei, w = gg.cosine_similarity_to_edge_index_weight(X, threshold=thr, keep_top_k=topk)
                assert ei.dtype == np.int64 and w.dtype == np.float64 and ei.shape[0] == 2
                assert np.all(ei[0] < ei[1])
                if _budget_binds(len(g["nodes"]), topk, ei_ref.shape[1]):
                    r = fast.tie_aware_compare(ei, w, ei_ref, w_ref)
                    assert r["ok"], r
                else:
                    assert_kf_threshold_parity(ei, w, ei_ref, w_ref, thr)
                    if topk is None and thr in (0.0, 0.8):
                        assert np.array_equal(ei, ei_ref)  # same emission order as the reference


def test_higher_stride_pruned_graphs_match_reference(gg):
    """Row f4: stride 2 / 3 / 5 (the last one non-overlapping) with edge_weight_threshold and the quantile
    keep_top_k, as the builders call it for the DB1.. graphs (kmer_ssgnn.py:215-226)."""
    z = np.load(os.path.join(GOLDEN, "stride_k4.npz"))
    reads = str(z["reads"]).split("\n")
    for stride in (2, 3, 5):
        _, pf = gg.weighted_directed_edges(reads, k=4, stride=stride, inlcudeUNK=False)
        for j in range(4):
            tag = f"s{stride}_{j}"
            thr = None if float(z[tag + "_thr"]) < 0 else float(z[tag + "_thr"])
            topk = None if float(z[tag + "_topk"]) < 0 else float(z[tag + "_topk"])
            g = gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False,
                                        edge_weight_threshold=thr, keep_top_k=topk)
            assert g.nodes() == z[tag + "_nodes"].tolist(), tag
            assert np.array_equal(g.edge_index.cpu().numpy(), z[tag + "_edge_index"]), tag
            assert np.array_equal(g.db.weight64.cpu().numpy(), z[tag + "_weight64"]), tag
            assert np.array_equal(g.weight.cpu().numpy(), z[tag + "_weight64"].astype(np.float32)), tag


def test_plain_dict_input_to_build(gg):
    g = load_golden(os.path.join(GOLDEN, "small_k4_ragged_N.npz"))
    pf = {(u, v): int(c) for u, v, c in zip(g["pair_u"], g["pair_v"], g["pair_count"])}
    graph = gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    assert graph.nodes() == g["nodes"].tolist()
    assert np.array_equal(graph.edge_index.cpu().numpy(), g["edge_index"])
    assert np.array_equal(graph.db.weight64.cpu().numpy(), g["weight64"])


# --------------------------------------------------------------------------- digests of full reference runs
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "digest_*.npz"))),
                         ids=lambda p: os.path.basename(p)[7:-4])
def test_matches_reference_digest(gg, path):
    z = np.load(path)
    n_reads, k, seed = int(z["n_reads"]), int(z["k"]), int(z["seed"])
    thr = float(z["threshold"])
    topk = None if float(z["keep_top_k"]) < 0 else float(z["keep_top_k"])
    small_ks = sorted(int(f.split("_")[1][1:]) for f in z.files if f.endswith("_E"))
    reads = fast.synthetic_reads(n_reads, 150, seed=seed)
    from genomic_gnn_b200 import core

    built = core.build_graph(reads, k, ",".join(map(str, small_ks)), edges_threshold=thr, edges_keep_top_k=topk)
    db = built.db
    assert db.num_nodes == int(z["num_nodes"]) and db.num_edges == int(z["num_db_edges"])
    assert _sha(db.node_codes.cpu().numpy()) == str(z["node_codes_sha"])
    assert _sha(db.edge_index.cpu().numpy()) == str(z["db_edge_index_sha"])
    assert _sha(db.weight.cpu().numpy()) == str(z["db_weight_sha"])
    for i, s in enumerate(small_ks):
        assert _sha(built.x[i].cpu().numpy()) == str(z[f"x_s{s}_sha_f32"])
        kf = built.kf[i]
        ei, w32 = kf.edge_index.cpu().numpy(), kf.weight.cpu().numpy()
        assert ei.shape[1] == int(z[f"kf_s{s}_E"])
        budget = int(db.num_nodes**2 * topk) if topk is not None else None
        if budget and ei.shape[1] == budget and kf.cutoff is not None:
            # top-n bound: everything strictly above the cut-off class must match the reference
            wmin = float(z[f"kf_s{s}_wmin"])
            above = kf.weight64.cpu().numpy() > wmin * (1 + 1e-9)
            assert int(above.sum()) == int(z[f"kf_s{s}_n_above_wmin"])
            keys, _ = sort_edges(ei[:, above], w32[above])
            assert _sha(keys) == str(z[f"kf_s{s}_above_set_sha"])
            rest = kf.weight64.cpu().numpy()[~above]
            assert np.all(np.abs(rest - wmin) <= 1e-9 * wmin)
            assert abs(kf.cutoff - wmin) <= 1e-12
        else:
            keys, ws = sort_edges(ei, w32)
            assert _sha(keys) == str(z[f"kf_s{s}_set_sha"])
            assert _sha(ws) == str(z[f"kf_s{s}_w32_sha"])


# --------------------------------------------------------------------------- oracle on seeded inputs
@pytest.mark.parametrize("k,n_reads,small_k,thr,topk", [
    (3, 10_000, "2", 0.0, None),          # BASELINE config 1
    (6, 20_000, "2,5", 0.0, 0.01),        # config 2 shape, top-n binding
    (7, 30_000, "2,5", 0.8, 0.01),        # config 3 shape, threshold binding
    (5, 3_000, "1,3", 0.55, None),
])
def test_pipeline_matches_oracle(gg, k, n_reads, small_k, thr, topk):
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=k)
    want = fast.build_graph(reads, k, small_k, edges_threshold=thr, edges_keep_top_k=topk)
    got = core.build_graph(reads, k, small_k, edges_threshold=thr, edges_keep_top_k=topk)
    assert np.array_equal(got.db.node_codes.cpu().numpy(), want["node_codes"])
    assert np.array_equal(got.db.edge_index.cpu().numpy(), want["edge_index"])
    assert np.array_equal(got.db.weight.cpu().numpy().view(np.int32), want["weight"].view(np.int32))
    for i in range(len(small_k.split(","))):
        assert np.array_equal(got.kf_counts[i].counts.cpu().numpy(), want[f"counts_{i}"])
        assert np.array_equal(got.x[i].cpu().numpy(), want[f"y_{i}"])
        # the oracle states the same deterministic tie policy, so even the top-n case is exact and ordered
        assert np.array_equal(got.kf[i].edge_index.cpu().numpy(), want[f"edge_index_KF{i}"])
        assert np.array_equal(got.kf[i].weight64.cpu().numpy(), want[f"edge_weight_KF{i}"])


def test_randomised_ragged_inputs_match_oracle(gg):
    """Forty seeded random inputs in the awkward corner of the input space -- reads of 0..45 bases, N runs, k from 2
    to 8 (reads shorter than k+1, windows bridged over N, single-pair graphs), string lists and fixed-length rows --
    through the whole device pipeline against the oracle: node order, De Bruijn edges, float32 weights, features,
    KF edges and float64 KF weights, all exact."""
    from genomic_gnn_b200 import core

    rng = np.random.default_rng(2024)
    done = 0
    for case in range(60):
        k = int(rng.integers(2, 9))
        n_reads = int(rng.integers(1, 70))
        p_n = float(rng.choice([0.0, 0.02, 0.15]))
        if case % 3 == 0:   # fixed-length rows (uint8 matrix input)
            length = int(rng.integers(k + 1, 46))
            arr = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n_reads, length))].copy()
            arr[rng.random(arr.shape) < p_n] = ord("N")
            reads, oracle_reads = arr, fast.reads_to_strings(arr)
        else:
            oracle_reads = []
            for _ in range(n_reads):
                ln = int(rng.integers(0, 46))
                r = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
                oracle_reads.append("".join("N" if rng.random() < p_n else ch for ch in r))
            reads = oracle_reads
        s = int(rng.integers(1, min(k, 3) + 1))
        thr = float(rng.choice([0.0, 0.5, 0.8]))
        topk = [None, 0.05][int(rng.integers(0, 2))]
        try:
            want = fast.build_graph(oracle_reads, k, str(s), edges_threshold=thr, edges_keep_top_k=topk)
        except (ValueError, IndexError):
            with pytest.raises((ValueError, IndexError)):      # no k-mer pair at all: both sides refuse
                core.build_graph(reads, k, str(s), edges_threshold=thr, edges_keep_top_k=topk)
            continue
        got = core.build_graph(reads, k, str(s), edges_threshold=thr, edges_keep_top_k=topk)
        tag = (case, k, s, thr, topk)
        assert np.array_equal(got.db.node_codes.cpu().numpy(), want["node_codes"]), tag
        assert np.array_equal(got.db.edge_index.cpu().numpy(), want["edge_index"]), tag
        assert np.array_equal(got.db.weight.cpu().numpy().view(np.int32), want["weight"].view(np.int32)), tag
        assert np.array_equal(got.x[0].cpu().numpy(), want["y_0"]), tag
        assert np.array_equal(got.kf[0].edge_index.cpu().numpy(), want["edge_index_KF0"]), tag
        assert np.array_equal(got.kf[0].weight64.cpu().numpy(), want["edge_weight_KF0"]), tag
        done += 1
    assert done >= 40


@pytest.mark.parametrize("k,s,thr", [(6, 3, 0.5), (6, 2, 0.8), (6, 2, 0.0), (5, 1, 0.9), (7, 2, 0.75), (3, 2, 0.3)])
def test_kf_impls_agree(gg, k, s, thr):
    """Narrow rows (D = 4, 16, 64): tensor path (rows padded to >= 32 bytes, 32B/64B swizzle), dp4a kernel and the
    generic slab kernel must emit identical edges in identical order."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(4000 if k < 7 else 900, 150, seed=11 + k)
    built = core.build_graph(reads, k, str(s), with_kf=False)
    kc = built.kf_counts[0]
    a = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dp4a")
    b = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="tcgen05")
    assert a.edge_index.shape[1] > 0
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight64, b.weight64)
    if 4**s == 64:
        c = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="naive")
        assert torch.equal(a.edge_index, c.edge_index) and torch.equal(a.weight64, c.weight64)
    else:  # duplicate-row classes (the default for <= 16 columns), whole and as a row-block shard
        c = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dedup")
        assert torch.equal(a.edge_index, c.edge_index) and torch.equal(a.weight64, c.weight64)
        nb = (built.db.num_nodes + 1023) // 1024
        lo, hi = (1, min(3, nb)) if nb >= 2 else (0, 1)
        a1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dp4a", iblock_begin=lo, iblock_end=hi)
        c1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dedup", iblock_begin=lo, iblock_end=hi)
        assert torch.equal(a1.edge_index, c1.edge_index) and torch.equal(a1.weight64, c1.weight64)
        with pytest.raises(ValueError):
            core.kf_edges(built.kf_counts[0].counts.repeat(1, 8), kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dedup")


@pytest.mark.parametrize("k,s,n_reads,thr", [
    (6, 4, 4000, 0.5),     # D = 256, N = 4096 (all 6-mers)
    (6, 5, 4000, 0.4),     # D = 1024, N = 4096
    (7, 5, 150, 0.3),      # D = 1024, ragged N (not every 7-mer occurs): partial strips, tiles and blocks
    (6, 4, 12, 0.0),       # tiny N, threshold 0: every pair sharing a 4-mer
    (8, 5, 3000, 0.7),     # D = 1024, N ~ 65k: many units per CTA, strip changes, deep ring reuse
])
def test_kf_tcgen05_matches_generic(gg, k, s, n_reads, thr):
    """TMA + tcgen05.mma.kind::i8 + TMEM epilogue vs the SIMT slab kernel: identical edges, order and weights."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=21 + k + s)
    built = core.build_graph(reads, k, str(s), with_kf=False)
    kc = built.kf_counts[0]
    b = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="naive")
    assert b.edge_index.shape[1] > 0
    for impl in ("tcgen05", "tcgen05x2"):   # single-CTA and CTA-pair (cta_group::2) kernels
        a = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl=impl)
        assert a.edge_index.shape == b.edge_index.shape, impl
        assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight64, b.weight64), impl
    # and a row-block shard of it
    nb = (built.db.num_nodes + 1023) // 1024
    if nb >= 3:
        b1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="naive", iblock_begin=1, iblock_end=3)
        for impl in ("tcgen05", "tcgen05x2"):
            a1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl=impl, iblock_begin=1, iblock_end=3)
            assert torch.equal(a1.edge_index, b1.edge_index), impl


@pytest.mark.parametrize("k,s,n_reads,thr", [
    (6, 4, 4000, 0.5),     # D = 256, N = 4096: a thread scans one tile
    (6, 5, 4000, 0.0),     # D = 1024, threshold 0: every pair that shares a column is an edge
    (7, 5, 150, 0.3),      # ragged N: partial last tile and row block
    (8, 5, 3000, 0.7),     # N ~ 65k: the BASELINE config 4 shape
    (8, 3, 3000, 0.9),     # D = 64: long posting lists (on request only)
])
def test_kf_sparse_join_matches_tensor_path(gg, k, s, n_reads, thr):
    """KF threshold edges of wide, sparse rows through the inverted-index join (impl="sparse") against the tensor / generic
    kernels: identical edges, order and weights, whole and as a shard."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=21 + k + s)
    built = core.build_graph(reads, k, str(s), with_kf=False)
    kc = built.kf_counts[0]
    ref_impl = "tcgen05x2" if 4**s >= 128 else "naive"
    a = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl=ref_impl)
    b = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="sparse", row_sum=kc.m)
    assert a.edge_index.shape[1] > 0
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight64, b.weight64)
    nb = (built.db.num_nodes + 1023) // 1024
    if nb >= 3:
        a1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl=ref_impl, iblock_begin=1, iblock_end=3)
        b1 = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="sparse", iblock_begin=1, iblock_end=3,
                           row_sum=kc.m)
        assert torch.equal(a1.edge_index, b1.edge_index) and torch.equal(a1.weight64, b1.weight64)
    # a record buffer that is too small: the join reports the exact count and is run again with room for it
    small = core.kf_emit(kc.counts, kc.sqnorm, thr, kc.m**2, impl="sparse", rec_capacity=64)
    assert small[1] == a.edge_index.shape[1] and small[0].shape[0] >= small[1] > 64
    # a binding top-n on the join's candidates: placed, histogrammed and selected as packed words
    top = 0.002
    ei, w = fast.kf_edges_exact(kc.counts.cpu().numpy(), thr, top) if built.db.num_nodes <= 4096 else (None, None)
    if ei is not None:
        cut = core.kf_edges(kc.counts, kc.sqnorm, thr, top, sq_max=kc.m**2, impl="sparse", row_sum=kc.m)
        assert np.array_equal(cut.edge_index.cpu().numpy(), ei) and np.array_equal(cut.weight64.cpu().numpy(), w)


@pytest.mark.parametrize("k,s,n_reads,thr", [
    (9, 2, 400_000, 0.93),   # N ~ 262k: 256 column blocks = four 64-block windows per class, shard across windows
    (8, 1, 200_000, 0.995),  # 4 columns: a few hundred classes with thousands of members (member sweeps, dense rows)
])
def test_kf_duplicate_row_classes_large(gg, k, s, n_reads, thr):
    """The duplicate-row path beyond one window / one member sweep, against the tensor-core scoring path."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=5 + k)
    built = core.build_graph(reads, k, str(s), with_kf=False)
    kc = built.kf_counts[0]
    nb = (built.db.num_nodes + 1023) // 1024
    for lo, hi in ((0, nb), (nb // 3, nb // 3 + max(nb // 4, 1))):
        a = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="tcgen05", iblock_begin=lo, iblock_end=hi)
        b = core.kf_edges(kc.counts, kc.sqnorm, thr, None, sq_max=kc.m**2, impl="dedup", iblock_begin=lo, iblock_end=hi)
        assert a.edge_index.shape[1] > 0
        assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight64, b.weight64)
        del a, b


def test_kf_row_block_shards_concatenate(gg):
    """Row-block shards (the multi-GPU unit) concatenate to the single-GPU result."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(5000, 150, seed=3)
    built = core.build_graph(reads, 6, "2", with_kf=False)   # N = 4096 -> 4 row blocks
    kc = built.kf_counts[0]
    full = core.kf_edges(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m**2)
    parts = [core.kf_edges(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m**2, iblock_begin=b, iblock_end=b + 1)
             for b in range(4)]
    assert torch.equal(torch.cat([p.edge_index for p in parts], dim=1), full.edge_index)
    assert torch.equal(torch.cat([p.weight64 for p in parts]), full.weight64)


def test_packed_shards_expand_like_a_multi_gpu_exchange(gg):
    """The multi-GPU KF exchange on one GPU: every row-block shard places its packed words into its segment of a
    padded all-ranks buffer, gg_kf_expand reads the segments as one array -- equals the unsharded edge list.  Both
    staging flavours (records from the tensor path, duplicate-row classes)."""
    import ctypes as C

    from genomic_gnn_b200 import _lib, core

    lib = _lib.load()
    reads = fast.synthetic_reads(5000, 150, seed=3)
    built = core.build_graph(reads, 6, "2", with_kf=False)   # N = 4096 -> 4 row blocks
    kc = built.kf_counts[0]
    full = core.kf_edges(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m**2)
    for impl in ("tcgen05", "dedup"):
        shards = [(0, 1), (1, 3), (3, 4)]
        stages = [core.kf_stage(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m**2, impl=impl, iblock_begin=a, iblock_end=b)
                  for a, b in shards]
        core.kf_resolve(stages)
        sizes = [st.staged.complete().n_recs for st in stages]
        stride = max(sizes)
        ordered = torch.zeros(len(shards) * stride, dtype=torch.int64, device="cuda")
        for r, st in enumerate(stages):
            st.staged.place(C.c_void_p(ordered.data_ptr() + 8 * r * stride))
        total = sum(sizes)
        ei = torch.empty((2, total), dtype=torch.int64, device="cuda")
        w = torch.empty(total, dtype=torch.float32, device="cuda")
        seg_first = (C.c_int64 * len(shards))(*[r * stride for r in range(len(shards))])
        seg_size = (C.c_int64 * len(shards))(*sizes)
        assert lib.gg_kf_expand(C.c_void_p(ordered.data_ptr()), len(shards), seg_first, seg_size,
                                C.c_void_p(kc.sqnorm.data_ptr()), C.c_void_p(ei[0].data_ptr()),
                                C.c_void_p(ei[1].data_ptr()), None, C.c_void_p(w.data_ptr()), None, None) == 0
        torch.cuda.synchronize()
        assert torch.equal(ei, full.edge_index) and torch.equal(w, full.weight), impl


def test_k10_million_nodes(gg):
    """BASELINE config 5's node count (k=10: 1,048,576 nodes, the most the KF kernels take) with thresholds that keep the
    edge lists in memory: the duplicate-row path and the tensor path agree edge for edge on the small_k=2 set
    (5.5e11 pairs), and sampled rows of the small_k=5 set (1.1e15 MACs through tcgen05) match numpy."""
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(2_000_000, 150, seed=0)
    built = core.build_graph(reads, 10, "2,5", with_kf=False)
    n = built.db.num_nodes
    assert n == 4**10 and built.db.num_edges == 4**11
    kc2, kc5 = built.kf_counts
    a = core.kf_edges(kc2.counts, kc2.sqnorm, 0.95, None, sq_max=kc2.m**2, impl="dedup")
    b = core.kf_edges(kc2.counts, kc2.sqnorm, 0.95, None, sq_max=kc2.m**2, impl="tcgen05")
    assert a.edge_index.shape[1] > 10_000_000
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.weight64, b.weight64)
    del a, b
    kf5 = core.kf_edges(kc5.counts, kc5.sqnorm, 0.8, None, sq_max=kc5.m**2)
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(n, 16, replace=False))
    c = kc5.counts.cpu().numpy()
    sq = kc5.sqnorm.cpu().numpy().astype(np.int64)
    a_rows = c[rows].astype(np.float32)                       # counts <= 6: float32 dots are exact
    dots = np.concatenate([a_rows @ c[lo:lo + 131072].T.astype(np.float32) for lo in range(0, n, 131072)], axis=1)
    cos = dots.astype(np.float64) / np.sqrt(sq[rows][:, None] * sq[None, :])
    ei = kf5.edge_index.cpu().numpy()
    assert np.all(ei[0] < ei[1])
    for r, row in zip(rows, cos):
        assert np.array_equal(np.nonzero((row > 0.8) & (np.arange(n) > r))[0], np.sort(ei[1][ei[0] == r])), r


# --------------------------------------------------------------------------- edge cases
def test_input_forms_agree(gg):
    from genomic_gnn_b200 import core

    arr = fast.synthetic_reads(500, 37, seed=5)
    strs = fast.reads_to_strings(arr)
    outs = []
    for inp in (arr, strs, torch.from_numpy(arr).cuda()):
        b = core.build_graph(inp, 4, "2", with_kf=False)
        outs.append((b.db.node_codes.cpu().numpy(), b.db.edge_index.cpu().numpy(), b.db.weight.cpu().numpy()))
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("k", [3, 6, 7, 8])
def test_count_paths_agree(gg, k):
    """The partitioned path, the shared-memory slices (packed stream) and the L2 RED path must give the same histogram, first occurrences and
    bridge records -- on ragged, N-ridden, separator-delimited input and on fixed-length rows."""
    from genomic_gnn_b200 import core

    rng = np.random.default_rng(100 + k)
    reads = []
    for _ in range(3000):
        ln = int(rng.integers(0, 200))
        r = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
        r = "".join(("N" if rng.random() < 0.01 else ch) for ch in r) + "N" * int(rng.integers(0, 4))
        reads.append(r)
    fixed = fast.synthetic_reads(5000, 151, seed=k)
    fixed[::41, 77] = ord("N")
    skewed = ["A" * 150] * 1500 + ["T" * 149] * 700 + ["ACGT" * 40] * 300  # every key of a tile in one bucket
    short_rows = fast.synthetic_reads(6000, 20, seed=k + 1)   # several read starts per 32 positions: general window path
    short_rows[::17, 9] = ord("N")
    odd_rows = fast.synthetic_reads(4000, 33, seed=k + 2)
    for inp in (reads, fixed, skewed, short_rows, odd_rows):
        stream = core.upload_reads(inp)
        b = core.count_kp1mers(stream, k, impl="l2")
        kb = sorted(map(tuple, b.bridges[: b.n_bridges].cpu().tolist()))
        for impl in ("part", "smem"):
            a = core.count_kp1mers(stream, k, impl=impl)
            assert torch.equal(a.hist, b.hist), impl
            assert torch.equal(a.first_pos, b.first_pos), impl
            assert a.n_bridges == b.n_bridges and a.n_nonzero == b.n_nonzero
            ka = sorted(map(tuple, a.bridges[: a.n_bridges].cpu().tolist()))
            assert ka == kb, impl
    for impl in ("part", "smem"):
        with pytest.raises(ValueError):
            core.count_kp1mers(core.upload_reads(["ACGTACGTACGTAAA"]), 10, impl=impl)
        with pytest.raises(ValueError):
            core.count_kp1mers(core.upload_reads(["ACGTXCGT"]), 3, impl=impl)


def test_count_pieces_merge_like_a_multi_gpu_build(gg):
    """gg_count_merge on one GPU: two shards counted into two exchange pieces (core.count_piece), laid side by side as
    an all-gather would, merged -- equals the count of the whole stream."""
    import ctypes as C

    from genomic_gnn_b200 import _lib, core

    lib = _lib.load()
    k = 7
    reads = fast.synthetic_reads(6000, 150, seed=12)
    reads[::29, 60] = ord("N")
    whole = core.count_kp1mers(core.upload_reads(reads), k)
    half = 3000
    pieces = []
    for lo, hi in ((0, half), (half, 6000)):
        piece = core.count_piece(k, torch.device("cuda"))
        core.count_kp1mers(core.upload_reads(reads[lo:hi], pos_base=lo * 150), k, sync=False, piece=piece)
        pieces.append(piece)
    everyone = torch.cat(pieces)
    bins = whole.hist.numel()
    hist = torch.empty(bins, dtype=torch.int32, device="cuda")
    first = torch.empty(bins, dtype=torch.int64, device="cuda")
    assert lib.gg_count_merge(C.c_void_p(everyone.data_ptr()), 2, pieces[0].numel(), k, C.c_void_p(hist.data_ptr()),
                              C.c_void_p(first.data_ptr()), None) == 0
    torch.cuda.synchronize()
    assert torch.equal(hist, whole.hist) and torch.equal(first, whole.first_pos)
    at = bins // 2 + bins                  # status words + stream end sit right behind first_pos (core.count_piece)
    tails = everyone.view(2, -1)[:, at: at + _lib.ST_WORDS + 1].cpu()
    assert int(tails[:, _lib.ST_BRIDGES].sum()) == whole.n_bridges and tails[:, -1].tolist() == [half * 150, 6000 * 150]


def test_host_to_host_build_matches_device_build(gg):
    from genomic_gnn_b200 import core

    arr = fast.synthetic_reads(3001, 150, seed=6)
    arr[::53, 20] = ord("N")
    dev = core.build_graph(arr, 6, "2,4", edges_threshold=0.8, edges_keep_top_k=0.01)
    for chunks in (1, 3, 8):
        host = core.build_graph_to_host(torch.from_numpy(arr), 6, "2,4", edges_threshold=0.8, edges_keep_top_k=0.01,
                                        chunks=chunks)
        assert not host.edge_index.is_cuda and host.edge_index.is_pinned()
        assert torch.equal(host.node_codes, dev.db.node_codes.cpu())
        assert torch.equal(host.edge_index, dev.db.edge_index.cpu()) and torch.equal(host.weight, dev.db.weight.cpu())
        for i in range(2):
            assert torch.equal(host.x[i], dev.x[i].cpu())
            assert torch.equal(host.kf_edge_index[i], dev.kf[i].edge_index.cpu())
            assert torch.equal(host.kf_weight[i], dev.kf[i].weight.cpu())
        assert host.h2d_bytes == arr.size
    # the download half on its own (what a multi-GPU build calls on the rank that owns the host copy)
    for expand in (True, False):
        host = core.graph_to_host(dev, host_expand=expand)
        assert host.x[1].is_pinned() and torch.equal(host.edge_index, dev.db.edge_index.cpu())
        for i in range(2):
            assert torch.equal(host.x[i], dev.x[i].cpu())
            assert torch.equal(host.kf_edge_index[i], dev.kf[i].edge_index.cpu())
            assert torch.equal(host.kf_weight[i], dev.kf[i].weight.cpu())


def test_degenerate_inputs(gg):
    with pytest.raises((ValueError, IndexError)):
        _, pf = gg.weighted_directed_edges([], k=3, inlcudeUNK=False)
        gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    with pytest.raises((ValueError, IndexError)):
        _, pf = gg.weighted_directed_edges(["AC", "", "NNNNNNN", "ACG"], k=3, inlcudeUNK=False)
        assert len(pf) == 0
        gg.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    with pytest.raises(ValueError):
        gg.weighted_directed_edges(["ACGTXACGT"], k=3, inlcudeUNK=False)
    with pytest.raises(ValueError):
        gg.weighted_directed_edges(["acgtacgt"], k=3, inlcudeUNK=False)
    with pytest.raises(NotImplementedError):
        gg.weighted_directed_edges(["ACGT"], k=3, inlcudeUNK=True)
    with pytest.raises(ValueError):
        gg.node_embedding_initial(_FakeGraph(), method="kmer_frequency", k=3)


class _FakeGraph:
    def number_of_nodes(self):
        return 0


def test_internal_n_bridge(gg):
    _, pf = gg.weighted_directed_edges(["ACGTNTTGA", "ACGTNTTGA", "CCCNNNCCCC"], k=3, inlcudeUNK=False)
    u, v, c, _ = fast.count_pairs(["ACGTNTTGA", "ACGTNTTGA", "CCCNNNCCCC"], 3)
    assert list(pf.keys()) == list(zip(fast.decode_kmers(u, 3), fast.decode_kmers(v, 3)))
    assert list(pf.values()) == c.tolist()
    assert pf[("CGT", "TTG")] == 2          # non-overlapping bridge
    assert pf[("CCC", "CCC")] == 2          # one overlapping bridge folded into the (k+1)-mer bin + one ordinary


def test_more_bridge_records_than_the_default_room(gg):
    """Every read holds an internal N run, so the one-read build (K1 + K2 + K3 behind one device read, assembled on the
    assumption of a stream without bridge records) has to fall back twice: bridge records exist, and there are more of
    them than the default 65,536 slots -- the stream is counted again with room for all."""
    from genomic_gnn_b200 import core

    rng = np.random.default_rng(17)
    bases = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for _ in range(70_000):
        a = bases[rng.integers(0, 4, 14)].tobytes().decode()
        b = bases[rng.integers(0, 4, 14)].tobytes().decode()
        reads.append(a + "N" * int(rng.integers(1, 3)) + b)
    want = fast.build_graph(reads, 4, "2", edges_threshold=0.9)
    got = core.build_graph(reads, 4, "2", edges_threshold=0.9)
    assert np.array_equal(got.db.node_codes.cpu().numpy(), want["node_codes"])
    assert np.array_equal(got.db.edge_index.cpu().numpy(), want["edge_index"])
    assert np.array_equal(got.db.weight.cpu().numpy(), want["weight"])
    assert np.array_equal(got.x[0].cpu().numpy(), want["y_0"])
    assert np.array_equal(got.kf[0].edge_index.cpu().numpy(), want["edge_index_KF0"])


def test_skewed_low_complexity_reads(gg):
    from genomic_gnn_b200 import core

    rng = np.random.default_rng(9)
    reads = ["A" * 150] * 400 + ["AC" * 75] * 300 + ["ACGT" * 37 + "AC"] * 300
    reads += fast.reads_to_strings(fast.synthetic_reads(200, 150, seed=2))
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    want = fast.build_graph(reads, 5, "2", edges_threshold=0.8)
    got = core.build_graph(reads, 5, "2", edges_threshold=0.8)
    assert np.array_equal(got.db.node_codes.cpu().numpy(), want["node_codes"])
    assert np.array_equal(got.db.edge_index.cpu().numpy(), want["edge_index"])
    assert np.array_equal(got.db.weight.cpu().numpy(), want["weight"])
    assert np.array_equal(got.kf[0].edge_index.cpu().numpy(), want["edge_index_KF0"])


# --------------------------------------------------------------------------- full-size properties (BASELINE config 4)
def test_full_size_properties_k8_1m(gg):
    from genomic_gnn_b200 import core

    n_reads, k, length = 1_000_000, 8, 150
    reads = fast.synthetic_reads(n_reads, length, seed=0)
    stream = core.upload_reads(reads)
    cr = core.count_kp1mers(stream, k)
    hist = cr.hist.cpu().numpy().view(np.uint32)
    assert int(hist.astype(np.int64).sum()) == n_reads * (length - k)       # every window counted once
    first = cr.first_pos.cpu().numpy().view(np.uint64)
    assert np.all(first[hist > 0] < n_reads * length) and len(np.unique(first[hist > 0])) == int((hist > 0).sum())
    # checksum of checksums: the first 100k reads alone reproduce the reference digest's DB graph order
    db = core.build_db(cr)
    assert db.num_nodes == 4**k and db.num_edges == 4 ** (k + 1)
    z = np.load(os.path.join(GOLDEN, "digest_cfg4_k8_100k.npz"))
    # node order is decided by first occurrences, all of which lie in the first 100k reads for this seed
    assert _sha(db.node_codes.cpu().numpy()) == str(z["node_codes_sha"])
    assert _sha(db.edge_index.cpu().numpy()) == str(z["db_edge_index_sha"])
    w = db.weight64.cpu().numpy()
    src = db.edge_index[0].cpu().numpy()
    assert np.all(np.diff(src) >= 0)                                        # grouped by source in node order
    assert np.allclose(np.bincount(src, weights=w), 1.0, rtol=0, atol=1e-12)
    kc = core.kf_counts(db.node_codes, k, 2)
    kf = core.kf_edges(kc.counts, kc.sqnorm, 0.8, 0.01, sq_max=kc.m**2)
    assert kf.edge_index.shape[1] == int(z["kf_s2_E"])                       # depends on the node set only
    keys, ws = sort_edges(kf.edge_index.cpu().numpy(), kf.weight.cpu().numpy())
    assert _sha(keys) == str(z["kf_s2_set_sha"]) and _sha(ws) == str(z["kf_s2_w32_sha"])
