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
