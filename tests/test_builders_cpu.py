"""Host logic of the builder glue that needs no GPU: the hop count of the NeighborLoader
(gnn_tasks_miniBatch/utils.py:48-64) and the PyG branch of networkx_to_dataloader
(gnn_tasks_miniBatch/dataloader.py:78-177), run against a recording stand-in for torch_geometric."""
import sys
import types

import numpy as np
import pytest
import torch

from genomic_gnn_b200 import builders


def test_count_edge_types_is_the_reference_rule():
    # utils.py:60-64: the most frequent edge type, not the number of layers
    assert builders.count_edge_types(["32_DB", "32_KF0", "1_KF0"]) == 2
    assert builders.count_edge_types(["32_DB", "32_KF0", "1_DB"]) == 2
    assert builders.count_edge_types(["32_DB", "32_DB", "32_DB", "1_DB"]) == 4
    assert builders.count_edge_types(["128_KF0", "1_KF0"]) == 2
    assert builders.count_edge_types(["32_DB"]) == 1
    # layer strings without an edge type (MLP / RGCN configurations): always 2
    assert builders.count_edge_types(["32", "32", "32"]) == 2
    assert builders.count_edge_types(["64"]) == 2
    with pytest.raises(IndexError):   # the reference indexes layers[0] as well
        builders.count_edge_types([])


class _Store(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, key, value):
        self[key] = value


class _FakeHetero:
    def __init__(self):
        self.stores = {}

    def __getitem__(self, key):
        return self.stores.setdefault(key if isinstance(key, str) else tuple(key), _Store())

    @property
    def edge_types(self):
        return [k for k in self.stores if isinstance(k, tuple)]


class _FakeData(_Store):
    pass


class _FakeNeighborLoader:
    calls = []

    def __init__(self, data, **kwargs):
        type(self).calls.append((data, kwargs))


@pytest.fixture
def fake_pyg(monkeypatch):
    tg = types.ModuleType("torch_geometric")
    tgd = types.ModuleType("torch_geometric.data")
    tgl = types.ModuleType("torch_geometric.loader")
    tgd.Data, tgd.HeteroData = _FakeData, _FakeHetero
    tgl.NeighborLoader = _FakeNeighborLoader
    tg.data, tg.loader = tgd, tgl
    for name, mod in (("torch_geometric", tg), ("torch_geometric.data", tgd), ("torch_geometric.loader", tgl)):
        monkeypatch.setitem(sys.modules, name, mod)
    _FakeNeighborLoader.calls = []
    return _FakeNeighborLoader


def _graph(n=7, e=11, seed=0):
    rng = np.random.default_rng(seed)
    return types.SimpleNamespace(edge_index=torch.from_numpy(rng.integers(0, n, size=(2, e))),
                                 weight=torch.from_numpy(rng.random(e).astype(np.float32)), num_nodes=n)


def test_mini_batch_pyg_branch_passes_the_reference_loader_arguments(fake_pyg):
    g = _graph()
    labels = np.random.default_rng(1).random((g.num_nodes, 16))
    strides = [_graph(seed=2)]
    loader, data = builders.networkx_to_dataloader_mini_batch(
        g, labels, labels_to_predict=[labels, labels[:, :4]], batch_size=256, kmer_frequency_loss_weight=0,
        layers_config=["32_DB", "32_KF0", "1_KF0"], ss_task="CL", graphs_strides=strides)
    assert isinstance(loader, _FakeNeighborLoader) and isinstance(data, _FakeHetero)
    (seen, kw), = fake_pyg.calls
    assert seen is data
    # dataloader.py:167-177
    assert kw["num_neighbors"] == [-1, -1]
    assert kw["batch_size"] == 256 and kw["shuffle"] is True and kw["directed"] is False and kw["pin_memory"] is True
    kind, nodes = kw["input_nodes"]
    assert kind == "node" and torch.equal(nodes, torch.arange(g.num_nodes))
    assert set(kw) == {"num_neighbors", "input_nodes", "batch_size", "pin_memory", "shuffle", "directed"}
    # dataloader.py:80-103: HeteroData fields and dtypes
    assert data["node"].x.dtype == torch.float32 and tuple(data["node"].x.shape) == (g.num_nodes, 16)
    assert torch.equal(data["node", "DB", "node"].edge_index, g.edge_index)
    assert torch.equal(data["node", "DB", "node"].edge_attr, g.weight)
    assert torch.equal(data["node", "DB1", "node"].edge_index, strides[0].edge_index)
    assert tuple(data["node"]["y_1"].shape) == (g.num_nodes, 4)
    assert ("node", "KF0", "node") not in data.stores        # kmer_frequency_loss_weight == 0: no KF sets


def test_mini_batch_pyg_branch_sampling_task_uses_connected_nodes(fake_pyg):
    g = types.SimpleNamespace(edge_index=torch.tensor([[0, 2, 3, 3], [2, 0, 3, 5]]), weight=torch.ones(4), num_nodes=8)
    builders.networkx_to_dataloader_mini_batch(g, np.zeros((8, 4)), batch_size=4, layers_config=["8"], ss_task="Sampling")
    (_, kw), = fake_pyg.calls
    assert kw["num_neighbors"] == [-1, -1]                  # no edge type in the layer strings -> 2 hops
    assert kw["input_nodes"][1].tolist() == [0, 2, 3, 5]      # dataloader.py:154-162: self loop (3, 3) ignored


def test_full_batch_pyg_branch_builds_a_data_object(fake_pyg):
    g = _graph()
    labels = np.random.default_rng(3).random((g.num_nodes, 16))
    loader, data = builders.networkx_to_dataloader_full_batch(g, labels, labels_to_predict=[labels])
    assert isinstance(data, _FakeData)
    assert torch.equal(data.edge_index, g.edge_index) and torch.equal(data.weight, g.weight)
    assert data.num_nodes == g.num_nodes and data.x.dtype == torch.float32 and len(data.y) == 1
    assert len(list(loader)) == 1                           # gnn_tasks/utils.py:41-45: dummy DataLoader([0])


def test_without_pyg_the_stand_ins_are_used():
    assert builders._pyg() is None                          # torch_geometric is absent from this image
    g = _graph()
    loader, data = builders.networkx_to_dataloader_mini_batch(g, np.zeros((g.num_nodes, 4)), batch_size=3,
                                                              layers_config=["8_DB"], ss_task="CL")
    assert isinstance(loader, builders.SeedNodeLoader) and isinstance(data, builders.HeteroGraphData)
    assert sorted(torch.cat(list(loader)).tolist()) == list(range(g.num_nodes))
