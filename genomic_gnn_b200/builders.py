"""Drop-in ``create_graphs`` for the reference's two builder classes.

    KmerSsgnn.create_graphs      src/representations/kmer_ssgnn.py:162-230
    KmerSsgnnMb.create_graphs    src/representations/kmer_ssgnn_miniBatch.py:160-285
    networkx_to_dataloader       src/representations/gnn_tasks/utils.py:8-51             (full batch)
    networkx_to_dataloader       src/representations/gnn_tasks_miniBatch/dataloader.py:21-179 (mini batch)

Both functions read the same attributes of ``self`` as the reference methods
and set the same ones (graph, kf_labels, dataloader, torch_graph, num_features),
so a maintainer can bind them with ``KmerSsgnn.create_graphs = create_graphs``
(INTEGRATION.md).  With torch_geometric installed the returned objects are real
``Data`` / ``HeteroData`` / ``NeighborLoader``; without it (this image) they are
the minimal stand-ins below with the same field names.
"""
from collections import Counter
from typing import List, Optional

import numpy as np
import torch

from . import api, core


def _pyg():
    """(Data, HeteroData, NeighborLoader) of torch_geometric, or None when it is not importable (this image).
    Looked up per call so that the tests can run the PyG branch against a recording stand-in in sys.modules."""
    try:
        from torch_geometric.data import Data, HeteroData
        from torch_geometric.loader import NeighborLoader
    except Exception:  # noqa: BLE001
        return None
    return Data, HeteroData, NeighborLoader


class GraphData(dict):
    """Stand-in for torch_geometric.data.Data: attribute and item access, .items()."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value

    def to(self, device):
        out = GraphData()
        for key, val in self.items():
            if isinstance(val, torch.Tensor):
                out[key] = val.to(device)
            elif isinstance(val, list):
                out[key] = [v.to(device) if isinstance(v, torch.Tensor) else v for v in val]
            else:
                out[key] = val
        return out


class HeteroGraphData:
    """Stand-in for torch_geometric.data.HeteroData with one node type:
    data["node"].x, data["node", "DB", "node"].edge_index / .edge_attr, ..."""

    def __init__(self):
        self._stores = {}

    @staticmethod
    def _key(key):
        return key if isinstance(key, str) else tuple(key)

    def __getitem__(self, key):
        return self._stores.setdefault(self._key(key), GraphData())

    @property
    def edge_types(self):
        return [k for k in self._stores if isinstance(k, tuple)]

    @property
    def node_types(self):
        return [k for k in self._stores if isinstance(k, str)]


class SeedNodeLoader:
    """Stand-in for NeighborLoader when torch_geometric is absent: yields shuffled
    seed-node batches (neighbour expansion itself is PyG machinery, out of scope)."""

    def __init__(self, data, input_nodes, batch_size, shuffle=True):
        self.data, self.input_nodes, self.batch_size, self.shuffle = data, input_nodes, batch_size, shuffle

    def __len__(self):
        return (self.input_nodes.numel() + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        idx = self.input_nodes[torch.randperm(self.input_nodes.numel())] if self.shuffle else self.input_nodes
        for i in range(0, idx.numel(), self.batch_size):
            yield idx[i : i + self.batch_size]


def _as_float_tensor(labels, device):
    if isinstance(labels, api.KFLabels):
        return labels.to_torch_float32().to(device)
    return torch.tensor(np.asarray(labels), dtype=torch.float).to(device)


def count_edge_types(layers_config):
    """gnn_tasks_miniBatch/utils.py:48-64: the number of hops the NeighborLoader samples = how often the most used
    edge type occurs among the layer strings ("<channels>_<edgeType>"), or 2 when the strings carry no edge type
    (MLP / RGCN configurations)."""
    layers = list(layers_config)
    if len(layers[0].split("_")) > 1:
        return max(Counter(layer.split("_")[1] for layer in layers).values())
    return 2


# --------------------------------------------------------------------------- a7
def networkx_to_dataloader_full_batch(graph, labels, labels_to_predict=None, batch_size=1, dummy_dataloader=True,
                                      device="cpu"):
    """gnn_tasks/utils.py:8-51 -> (dataloader, Data)."""
    pyg = _pyg()
    data = pyg[0]() if pyg else GraphData()
    data.edge_index = graph.edge_index.to(device)
    data.weight = graph.weight.to(device)
    data.num_nodes = graph.num_nodes
    data.x = _as_float_tensor(labels, device)
    data.y = [_as_float_tensor(l, device) for l in (labels_to_predict or [])]
    loader = torch.utils.data.DataLoader([0] if dummy_dataloader else [data], batch_size=1 if dummy_dataloader else batch_size,
                                         num_workers=0, collate_fn=None if dummy_dataloader else (lambda b: b))
    return loader, data


# --------------------------------------------------------------------------- a8
def networkx_to_dataloader_mini_batch(graph, labels, labels_to_predict=None, batch_size=4096,
                                      kmer_frequency_loss_weight: float = 0, edit_distance_loss_weight: float = 0,
                                      kmer_freq_labels: Optional[list] = None, k: int = 3,
                                      edges_threshold: float = 0.0, edges_keep_top_k: Optional[float] = None,
                                      layers_config: Optional[list] = None, ss_task: Optional[str] = None,
                                      graphs_strides: Optional[list] = None, faiss_ann: bool = False, device="cpu",
                                      kf_impl="auto", faiss_distance: str = "L2", **_faiss_kwargs):
    """gnn_tasks_miniBatch/dataloader.py:21-179 -> (loader, HeteroData).  With ``faiss_ann`` the KF edges are every
    node's int(edges_keep_top_k * N) nearest nodes (dataloader.py:108-121), found exactly instead of through an
    approximate FAISS index (index type / nlist / nprobe / m / nbits are accepted and ignored)."""
    if edit_distance_loss_weight != 0:
        raise NotImplementedError("edit-distance (ED) edges are out of scope (SURVEY.md section 2, row 8)")
    pyg = _pyg()
    data = pyg[1]() if pyg else HeteroGraphData()
    data["node"].x = _as_float_tensor(labels, device)
    data["node", "DB", "node"].edge_index = graph.edge_index.to(device)
    data["node", "DB", "node"].edge_attr = graph.weight.to(device)
    for i, g in enumerate(graphs_strides or []):  # dataloader.py:83-97
        data["node", f"DB{i + 1}", "node"].edge_index = g.edge_index.to(device)
        data["node", f"DB{i + 1}", "node"].edge_attr = g.weight.to(device)
    for i, lab in enumerate(labels_to_predict or []):
        data["node"][f"y_{i}"] = _as_float_tensor(lab, device)
    if kmer_frequency_loss_weight != 0:
        for i, lab in enumerate(kmer_freq_labels or []):
            if faiss_ann:
                if not edges_keep_top_k:   # parsers.py:566-573 asserts the same before training starts
                    raise ValueError("Faiss ANN only supported with keep_top_k.")
                k_nn = int(edges_keep_top_k * len(lab))
                if isinstance(lab, api.KFLabels):
                    res = core.kf_knn_edges(lab.kc.counts, lab.kc.sqnorm, k_nn, faiss_distance)
                    ei, w = res.edge_index, res.weight
                else:
                    ei_np, w_np = api.kf_faiss_to_edge_index_weight(np.vstack(lab), k=k_nn, distance=faiss_distance)
                    ei, w = torch.tensor(ei_np, dtype=torch.long), torch.tensor(w_np, dtype=torch.float)
            elif isinstance(lab, api.KFLabels):
                kc = lab.kc
                res = core.kf_edges(kc.counts, kc.sqnorm, float(edges_threshold), edges_keep_top_k,
                                    sq_max=kc.m * kc.m, impl=kf_impl)
                ei, w = res.edge_index, res.weight
            else:
                ei_np, w_np = api.cosine_similarity_to_edge_index_weight(np.vstack(lab), threshold=edges_threshold,
                                                                         keep_top_k=edges_keep_top_k)
                ei, w = torch.tensor(ei_np, dtype=torch.long), torch.tensor(w_np, dtype=torch.float)
            data["node", f"KF{i}", "node"].edge_index = ei.to(device)
            data["node", f"KF{i}", "node"].edge_attr = w.to(device)
    n = graph.num_nodes
    if ss_task == "Sampling":  # dataloader.py:154-162 (never true for the value the builders pass, "CL")
        allidx = torch.cat([data[rel].edge_index for rel in data.edge_types], dim=1)
        input_nodes = torch.unique(allidx[:, allidx[0] != allidx[1]])
    else:
        input_nodes = torch.arange(n)
    if pyg:  # dataloader.py:167-177
        loader = pyg[2](data, num_neighbors=[-1] * count_edge_types(layers_config), input_nodes=("node", input_nodes),
                        batch_size=batch_size, pin_memory=True, shuffle=True, directed=False)
    else:
        loader = SeedNodeLoader(data, input_nodes, batch_size, shuffle=True)
    return loader, data


# --------------------------------------------------------------------------- builder methods
def _stride_graphs(self, k):
    """kmer_ssgnn.py:215-226 / kmer_ssgnn_miniBatch.py:201-217: one extra DB graph per additional stride, pruned by
    edge_weight_threshold = 1/4^k and the quantile keep_top_k.  Each has its own node numbering (from_networkx of the
    new graph), as in the reference."""
    out = []
    for stride in list(getattr(self, "strides", [1]))[1:]:
        _, pf = api.weighted_directed_edges(self.ssgnn_dataset, k=k, stride=int(stride), inlcudeUNK=False)
        out.append(api.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=self.create_all_kmers,
                                            edge_weight_threshold=1.0 / float(4**k),
                                            keep_top_k=getattr(self, "edges_keep_top_k", None)))
    return out


def _common_graph(self, k):
    strides = list(getattr(self, "strides", [1]))
    _, pair_frequency = api.weighted_directed_edges(self.ssgnn_dataset, k=k, stride=int(strides[0]), inlcudeUNK=False,
                                                    disable_tqdm=getattr(self, "disable_tqdm", True))
    self.graph = api.build_deBruijn_graph(pair_frequency, normalise=True, remove_N=True,
                                          create_all_kmers=self.create_all_kmers,
                                          disable_tqdm=getattr(self, "disable_tqdm", True))
    self.kf_labels = [api.node_embedding_initial(self.graph, method="kmer_frequency", k=k, small_k=int(s))
                      for s in str(self.small_k).split(",")]


def _fused_host_build(self, k, with_kf):
    """The builders' usual case in ONE pipelined pass (core.build_graph_to_host): list[str] reads marshalled natively
    under their own upload, kernels back to back without intermediate Python objects, every tensor copied to pinned
    host memory under the kernels that follow it.  Usual case = single stride, k-mer-frequency labels, exact KF edges,
    CPU-side loaders (what kmer_ssgnn.py:162-213 / kmer_ssgnn_miniBatch.py:160-199 run with the default flags).
    Returns None when the configuration needs the step-by-step path."""
    if (len(list(getattr(self, "strides", [1]))) != 1 or int(list(getattr(self, "strides", [1]))[0]) != 1
            or getattr(self, "representation_ss_initial_labels", "kmer_frequency") != "kmer_frequency"
            or getattr(self, "faiss_ann", False) or str(getattr(self, "graph_device", "cpu")) != "cpu"
            or getattr(self, "gg_stepwise", False)   # opt-out: run the four API calls one by one
            or not isinstance(self.ssgnn_dataset, (list, tuple)) or len(self.ssgnn_dataset) == 0):
        return None
    host = core.build_graph_to_host(self.ssgnn_dataset, k, str(self.small_k), create_all_kmers=self.create_all_kmers,
                                    edges_threshold=float(getattr(self, "edges_threshold", 0.0) or 0.0),
                                    edges_keep_top_k=getattr(self, "edges_keep_top_k", None), with_kf=with_kf)
    self.graph = api.KmerGraph(host.db)
    self.kf_labels = [api.KFLabels(kc) for kc in host.kf_counts]
    return host


def create_graphs(self, k: int):
    """Full-batch builder: KmerSsgnn.create_graphs (kmer_ssgnn.py:162-230)."""
    host = _fused_host_build(self, k, with_kf=False)
    if host is not None:
        pyg = _pyg()
        data = pyg[0]() if pyg else GraphData()
        data.edge_index, data.weight, data.num_nodes = host.edge_index, host.weight, host.num_nodes
        data.x, data.y = host.x[0], list(host.x)
        self.dataloader = torch.utils.data.DataLoader([0], batch_size=1, num_workers=0)
        self.torch_graph = data
        self.num_features = 4 ** int(str(self.small_k)[0])  # sic: first character (kmer_ssgnn.py:213)
        return
    _common_graph(self, k)
    device = getattr(self, "graph_device", "cpu")
    if getattr(self, "representation_ss_initial_labels", "kmer_frequency") == "one_hot":
        onehot = api.node_embedding_initial(self.graph, method="onehot", k=k)
        self.dataloader, self.torch_graph = networkx_to_dataloader_full_batch(
            self.graph, onehot, labels_to_predict=self.kf_labels, device=device)
        self.num_features = 4**k
    else:
        self.dataloader, self.torch_graph = networkx_to_dataloader_full_batch(
            self.graph, self.kf_labels[0], labels_to_predict=self.kf_labels, device=device)
        self.num_features = 4 ** int(str(self.small_k)[0])  # sic: first character (kmer_ssgnn.py:213)
    for i, g in enumerate(_stride_graphs(self, k)):  # attach_higher_strides_to_torch_graph, gnn_utils.py:496-511
        self.torch_graph["edge_index_" + str(i + 1)] = g.edge_index.to(device)
        self.torch_graph["edge_weight_" + str(i + 1)] = g.weight.to(device)


def create_graphs_mini_batch(self, k: int):
    """Mini-batch builder: KmerSsgnnMb.create_graphs (kmer_ssgnn_miniBatch.py:160-285)."""
    if getattr(self, "edit_distance_loss_weight", 0) != 0:
        raise NotImplementedError("edit-distance (ED) edges are out of scope (SURVEY.md section 2, row 8)")
    with_kf = getattr(self, "kmer_frequency_loss_weight", 0) != 0
    host = _fused_host_build(self, k, with_kf=with_kf)
    if host is not None:
        pyg = _pyg()
        data = pyg[1]() if pyg else HeteroGraphData()
        data["node"].x = host.x[0]
        data["node", "DB", "node"].edge_index = host.edge_index
        data["node", "DB", "node"].edge_attr = host.weight
        for i, x in enumerate(host.x):
            data["node"][f"y_{i}"] = x
        for i, (ei, w) in enumerate(zip(host.kf_edge_index, host.kf_weight)):
            data["node", f"KF{i}", "node"].edge_index = ei
            data["node", f"KF{i}", "node"].edge_attr = w
        input_nodes = torch.arange(host.num_nodes)   # dataloader.py:154-165: ss_task is "CL" here, never "Sampling"
        layers = [*self.layers_config, "1_" + self.representation_ss_last_layer_edge_type]
        if getattr(self, "ss_task", None) == "Sampling":
            allidx = torch.cat([data[rel].edge_index for rel in data.edge_types], dim=1)
            input_nodes = torch.unique(allidx[:, allidx[0] != allidx[1]])
        if pyg:
            loader = pyg[2](data, num_neighbors=[-1] * count_edge_types(layers), input_nodes=("node", input_nodes),
                            batch_size=self.batch_size, pin_memory=True, shuffle=True, directed=False)
        else:
            loader = SeedNodeLoader(data, input_nodes, self.batch_size, shuffle=True)
        self.dataloader, self.torch_graph = loader, data
        self.num_features = 4 ** int(str(self.small_k)[0])
        self.build_info = {"marshal_ms": host.marshal_ms, "h2d_bytes": host.h2d_bytes, "d2h_bytes": host.d2h_bytes}
        return
    _common_graph(self, k)
    device = getattr(self, "graph_device", "cpu")
    one_hot = getattr(self, "representation_ss_initial_labels", "kmer_frequency") == "one_hot"
    labels = api.node_embedding_initial(self.graph, method="onehot", k=k) if one_hot else self.kf_labels[0]
    self.dataloader, self.torch_graph = networkx_to_dataloader_mini_batch(
        self.graph, labels, labels_to_predict=self.kf_labels, batch_size=self.batch_size,
        kmer_frequency_loss_weight=self.kmer_frequency_loss_weight,
        edit_distance_loss_weight=self.edit_distance_loss_weight, kmer_freq_labels=self.kf_labels, k=k,
        edges_threshold=self.edges_threshold, edges_keep_top_k=self.edges_keep_top_k,
        layers_config=[*self.layers_config, "1_" + self.representation_ss_last_layer_edge_type],
        ss_task=self.ss_task, graphs_strides=_stride_graphs(self, k), faiss_ann=getattr(self, "faiss_ann", False),
        faiss_distance=getattr(self, "faiss_distance", "L2"), device=device)
    self.num_features = 4**k if one_hot else 4 ** int(str(self.small_k)[0])


def kf_edges_config(kf_labels: List[api.KFLabels], edges_threshold=0.0, edges_keep_top_k=None, device="cpu"):
    """What SamplingGeneral.__init__ (gnn_tasks/sampling_task.py:128-161) derives from
    the labels: edges_config entries edge_index_KF{i} / edge_weight_KF{i} and the
    sampling probabilities p_vectors_KF (weights / weights.sum())."""
    cfg, p_vectors, sample_index = {}, [], []
    for i, lab in enumerate(kf_labels):
        kc = lab.kc
        res = core.kf_edges(kc.counts, kc.sqnorm, float(edges_threshold), edges_keep_top_k, sq_max=kc.m * kc.m)
        cfg[f"edge_index_KF{i}"] = res.edge_index.to(device)
        cfg[f"edge_weight_KF{i}"] = res.weight.to(device)
        w = res.weight64.cpu().numpy()
        p_vectors.append(w / w.sum() if w.size else w)
        sample_index.append(res.edge_index.t().cpu().numpy())
    return cfg, p_vectors, sample_index
