"""Drop-in builder methods (kmer_ssgnn.py:162-230, kmer_ssgnn_miniBatch.py:160-285) on a stand-in for `self`."""
import types

import numpy as np
import pytest
import torch

from oracle import fast, faithful

pytestmark = pytest.mark.gpu


def _self(reads, **kw):
    base = dict(ssgnn_dataset=reads, strides=[1], small_k="2,3", create_all_kmers=False, disable_tqdm=True,
                representation_ss_initial_labels="kmer_frequency", batch_size=64, kmer_frequency_loss_weight=1.0,
                edit_distance_loss_weight=0.0, edges_threshold=0.8, edges_keep_top_k=0.01, layers_config=["32_DB", "32_KF0"],
                representation_ss_last_layer_edge_type="KF0", ss_task="CL", faiss_ann=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _reference(reads, k, small_k, thr, topk):
    return faithful.build_reference_graph(reads, k, small_k, edges_threshold=thr, edges_keep_top_k=topk)


def test_full_batch_create_graphs():
    import genomic_gnn_b200 as gg

    reads = fast.reads_to_strings(fast.synthetic_reads(400, 150, seed=8))
    me = _self(reads)
    gg.create_graphs(me, 4)
    ref = _reference(reads, 4, "2,3", 0.8, None)
    data = me.torch_graph
    assert me.graph.nodes() == ref["nodes"] and me.graph.number_of_nodes() == ref["num_nodes"]
    assert me.num_features == 16                                   # 4 ** int("2,3"[0])
    assert data.edge_index.dtype == torch.int64 and data.weight.dtype == torch.float32 and data.x.dtype == torch.float32
    assert not data.edge_index.is_cuda                              # handed to the reference classes on the CPU
    assert np.array_equal(data.edge_index.numpy(), ref["edge_index"])
    assert np.array_equal(data.weight.numpy(), ref["weight"])
    assert np.array_equal(data.x.numpy(), ref["x"])
    assert len(data.y) == 2 and all(np.array_equal(a.numpy(), b) for a, b in zip(data.y, ref["y"]))
    assert data.num_nodes == ref["num_nodes"]
    assert len(me.kf_labels) == 2 and np.array_equal(np.asarray(me.kf_labels[1]), ref["kf_labels"][1])
    assert list(me.dataloader) == [torch.tensor([0])] or len(list(me.dataloader)) == 1
    cfg, p_vectors, sample_index = gg.kf_edges_config(me.kf_labels, 0.8, None)
    for i in range(2):
        assert np.array_equal(cfg[f"edge_index_KF{i}"].numpy(), ref[f"edge_index_KF{i}"])
        assert np.array_equal(cfg[f"edge_weight_KF{i}"].numpy(), ref[f"edge_weight_KF{i}"].astype(np.float32))
        w = ref[f"edge_weight_KF{i}"]
        assert np.allclose(p_vectors[i], w / w.sum(), rtol=1e-12, atol=0)
        assert np.array_equal(sample_index[i], ref[f"edge_index_KF{i}"].T)


def test_full_batch_one_hot_labels():
    import genomic_gnn_b200 as gg

    reads = fast.reads_to_strings(fast.synthetic_reads(100, 60, seed=2))
    me = _self(reads, representation_ss_initial_labels="one_hot", small_k="2")
    gg.create_graphs(me, 3)
    assert me.num_features == 4**3
    assert torch.equal(me.torch_graph.x, torch.eye(me.graph.number_of_nodes()))


def test_mini_batch_create_graphs():
    import genomic_gnn_b200 as gg

    reads = fast.reads_to_strings(fast.synthetic_reads(400, 150, seed=9))
    me = _self(reads, edges_keep_top_k=None, edges_threshold=0.8)
    gg.create_graphs_mini_batch(me, 4)
    ref = _reference(reads, 4, "2,3", 0.8, None)
    data = me.torch_graph
    assert np.array_equal(data["node"].x.numpy(), ref["x"])
    assert np.array_equal(data["node", "DB", "node"].edge_index.numpy(), ref["edge_index"])
    assert np.array_equal(data["node", "DB", "node"].edge_attr.numpy(), ref["weight"])
    for i in range(2):
        assert np.array_equal(data["node"][f"y_{i}"].numpy(), ref["y"][i])
        assert np.array_equal(data["node", f"KF{i}", "node"].edge_index.numpy(), ref[f"edge_index_KF{i}"])
        assert np.array_equal(data["node", f"KF{i}", "node"].edge_attr.numpy(), ref[f"edge_weight_KF{i}"].astype(np.float32))
        assert data["node", f"KF{i}", "node"].edge_index.dtype == torch.int64
    seeds = torch.cat(list(me.dataloader))
    assert sorted(seeds.tolist()) == list(range(ref["num_nodes"]))
    assert me.num_features == 16


def test_higher_stride_graphs_are_attached():
    """kmer_ssgnn.py:215-230 / dataloader.py:83-97: DB1.. edge sets, each numbered by its own graph."""
    import genomic_gnn_b200 as gg

    reads = fast.reads_to_strings(fast.synthetic_reads(300, 120, seed=4))
    k = 4
    u, v, c, _ = fast.count_pairs(reads, k, 3)
    _, ei_ref, w_ref = fast.build_db(u, v, c, k, True, False, 1.0 / 4**k, 0.4)
    me = _self(reads, strides=[1, 3], edges_keep_top_k=0.4, edges_threshold=0.8)
    gg.create_graphs(me, k)
    assert np.array_equal(me.torch_graph["edge_index_1"].numpy(), ei_ref)
    assert np.array_equal(me.torch_graph["edge_weight_1"].numpy(), w_ref.astype(np.float32))
    mb = _self(reads, strides=[1, 3], edges_keep_top_k=0.4, edges_threshold=0.8)
    gg.create_graphs_mini_batch(mb, k)
    assert np.array_equal(mb.torch_graph["node", "DB1", "node"].edge_index.numpy(), ei_ref)
    assert np.array_equal(mb.torch_graph["node", "DB1", "node"].edge_attr.numpy(), w_ref.astype(np.float32))


def test_unsupported_options_fail_loudly():
    import genomic_gnn_b200 as gg

    reads = fast.reads_to_strings(fast.synthetic_reads(50, 60, seed=1))
    with pytest.raises(NotImplementedError):
        gg.create_graphs_mini_batch(_self(reads, edit_distance_loss_weight=1.0), 3)


# --------------------------------------------------------------------------- BASELINE-size builder runs against reference digests
def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("digest,batch", [("digest_cfg3_k7_100k.npz", 1024), ("digest_cfg4_k8_100k.npz", 256)])
def test_mini_batch_builder_at_baseline_sizes_matches_reference_digest(digest, batch):
    """create_graphs(self, k) of KmerSsgnnMb as the reference calls it -- ``ssgnn_dataset`` a list of 100,000 read
    strings, config 3 / config 4 flags (k=7 / 8, small_k=2,5, threshold 0.8, top_k 0.01, batch 1024 / 256) -- against
    digests of the REAL reference's tensors (oracle/make_golden.py): node order, DB edges and weights, float32
    features, every KF edge set and its float32 weights."""
    import os

    import genomic_gnn_b200 as gg
    from ._util import GOLDEN, sort_edges

    z = np.load(os.path.join(GOLDEN, digest))
    k, n_reads = int(z["k"]), int(z["n_reads"])
    reads = fast.reads_to_strings(fast.synthetic_reads(n_reads, 150, seed=int(z["seed"])))
    me = _self(reads, small_k="2,5", batch_size=batch, edges_threshold=float(z["threshold"]),
               edges_keep_top_k=float(z["keep_top_k"]), layers_config=["32_KF0"])
    gg.create_graphs_mini_batch(me, k)
    data = me.torch_graph
    assert me.graph.number_of_nodes() == int(z["num_nodes"]) and me.num_features == 16
    assert _sha(gg.api.encode_kmers(me.graph.nodes(), k)) == str(z["node_codes_sha"])
    db = data["node", "DB", "node"]
    assert not db.edge_index.is_cuda and db.edge_index.dtype == torch.int64 and db.edge_attr.dtype == torch.float32
    assert _sha(db.edge_index.numpy()) == str(z["db_edge_index_sha"]) and _sha(db.edge_attr.numpy()) == str(z["db_weight_sha"])
    assert _sha(data["node"].x.numpy()) == str(z["x_s2_sha_f32"])
    for i, s in enumerate((2, 5)):
        assert _sha(data["node"][f"y_{i}"].numpy()) == str(z[f"x_s{s}_sha_f32"])
        kf = data["node", f"KF{i}", "node"]
        assert kf.edge_index.shape[1] == int(z[f"kf_s{s}_E"])            # threshold-bound at these configs
        keys, ws = sort_edges(kf.edge_index.numpy(), kf.edge_attr.numpy())
        assert _sha(keys) == str(z[f"kf_s{s}_set_sha"]) and _sha(ws) == str(z[f"kf_s{s}_w32_sha"])
    assert len(me.dataloader) == -(-int(z["num_nodes"]) // batch)
    assert me.build_info["h2d_bytes"] == n_reads * 150 and me.build_info["marshal_ms"] > 0
    # kf_labels stay usable as the reference's float64 matrices (np.vstack at sampling_task.py:146-150)
    assert np.vstack(me.kf_labels[0]).shape == (int(z["num_nodes"]), 16)


def test_fused_builder_equals_stepwise_builder_on_ragged_reads():
    """The one-pass host build (list[str] -> native packing -> pipelined build) and the step-by-step API calls give the
    same builder attributes, on reads of unequal length with N runs (the '\\n'-separated stream layout)."""
    import genomic_gnn_b200 as gg
    from genomic_gnn_b200 import builders

    rng = np.random.default_rng(4)
    reads = []
    for r in fast.reads_to_strings(fast.synthetic_reads(600, 120, seed=6)):
        r = r[: int(rng.integers(0, 121))]
        reads.append("".join("N" if rng.random() < 0.02 else ch for ch in r))
    fused, step = _self(reads, small_k="2,3", edges_keep_top_k=None), _self(tuple(reads), small_k="2,3", edges_keep_top_k=None)
    gg.create_graphs_mini_batch(fused, 4)
    assert hasattr(fused, "build_info")
    step.gg_stepwise = True                                    # the four reference-signature calls, one by one
    assert builders._fused_host_build(step, 4, True) is None
    gg.create_graphs_mini_batch(step, 4)
    assert fused.graph.nodes() == step.graph.nodes()
    a, b = fused.torch_graph, step.torch_graph
    for rel in (("node", "DB", "node"), ("node", "KF0", "node"), ("node", "KF1", "node")):
        assert torch.equal(a[rel].edge_index, b[rel].edge_index) and torch.equal(a[rel].edge_attr, b[rel].edge_attr)
    assert torch.equal(a["node"].x, b["node"].x) and torch.equal(a["node"]["y_1"], b["node"]["y_1"])
    with pytest.raises(ValueError):
        gg.create_graphs_mini_batch(_self(["ACGT", 7]), 3)


def test_shared_host_arena_delivery_single_rank():
    """hostshare.deliver_graph with one rank: every tensor of a device build lands in the page-locked shared-memory
    arena (the multi-rank flavour, where every rank writes its own shard, runs in tests/test_multi_gpu.py)."""
    from genomic_gnn_b200 import core, hostshare

    reads = fast.synthetic_reads(5000, 150, seed=3)
    built = core.build_graph(reads, 6, "2,4", edges_threshold=0.8, edges_keep_top_k=0.01)
    arena = hostshare.SharedHostArena(None, hostshare.arena_bytes_for(built.db.num_nodes, 6, [2, 4],
                                                                       [kf.n_edges for kf in built.kf]))
    try:
        host = hostshare.deliver_graph(built, arena, None)
        assert torch.equal(host.node_codes, built.db.node_codes.cpu()) and torch.equal(host.edge_index, built.db.edge_index.cpu())
        assert torch.equal(host.weight, built.db.weight.cpu())
        for i, kf in enumerate(built.kf):
            assert torch.equal(host.x[i], built.x[i].cpu())
            assert torch.equal(host.kf_edge_index[i], kf.edge_index.cpu()) and torch.equal(host.kf_weight[i], kf.weight.cpu())
        assert host.d2h_bytes > 0 and not host.edge_index.is_cuda
        del host
    finally:
        arena.close()
