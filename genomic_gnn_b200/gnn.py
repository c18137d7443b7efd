"""Row f1: the consumer of the graph -- a PyG-free GCN encoder, the contrastive sampling task that trains it, and the
frozen-embedding edit-distance regressor downstream -- so that "downstream results unchanged" can be run, not argued.

Mirrors (reference paths):
  * ``GCN`` / ``interpret_layers_config``            gnn_common/gnn_models.py:8-100 (torch_geometric GCNConv 2.3.1:
    gcn_norm with remaining self loops of weight 1, symmetric degree normalisation on the target side, bias after
    aggregation; glorot weights, zero bias)
  * ``SamplingTask``                                 gnn_tasks/sampling_task.py:236-382 (positive pairs from biased
    random walks and from the KF edges in proportion to their weights, uniform negatives, -logsigmoid losses, Adam)
  * ``EditDistanceMLP`` / ``hyperbolic_distance``    downstream_tasks/edit_distance_models/mlp.py:11-235, distances.py:20-28

torch supplies autograd, dense GEMMs and the optimiser; the sparse aggregation is ``gg_spmm_csr_f32`` (csrc/gg_spmm.cu),
the positive pairs come from ``gg_rw_pairs``.  torch_geometric and pytorch_lightning are not needed.  The reference's
trainer cannot be imported here (both packages are absent), so this row's parity is pinned on the formulas only:
tests/test_gpu_gnn.py checks the aggregation kernel against dense float64 algebra (1e-5), the layer against a dense
restatement of GCNConv, and the whole run -- encoder training, embeddings, downstream regression -- for equal results on
the CUDA-built graph and on the oracle's graph (equal up to the float atomics of torch's own index backward: two runs of
the same program agree to ~1e-7 per step, not bit for bit).
"""
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr
from .core import _require_cuda, _stream
from .sampler import sample_biased_random_walks


def interpret_layers_config(layers_config: List[str]):
    """gnn_models.py:8-27: "32_DB" -> {"channels": 32, "edge_index": "edge_index_DB", "edge_weight": "edge_weight_DB"}."""
    out = []
    for config in layers_config:
        channels, edge = config.split("_")
        out.append({"channels": int(channels), "edge_index": f"edge_index_{edge}", "edge_weight": f"edge_weight_{edge}"})
    return out


# --------------------------------------------------------------------------- normalised adjacency
class NormAdj:
    """GCNConv's normalised adjacency of one edge set, as CSR by target (forward) and by source (backward)."""

    def __init__(self, edge_index: torch.Tensor, edge_weight: Optional[torch.Tensor], num_nodes: int):
        dev = edge_index.device
        row, col = edge_index[0], edge_index[1]
        w = torch.ones(row.numel(), dtype=torch.float32, device=dev) if edge_weight is None else edge_weight.float()
        # add_remaining_self_loops(fill_value=1): existing self loops keep their weight, every other node gets one of 1
        loop = row == col
        loop_w = torch.ones(num_nodes, dtype=torch.float32, device=dev)
        if bool(loop.any()):   # several self loops on one node: the last one in edge order wins (PyG's indexed assignment)
            pos = torch.nonzero(loop).reshape(-1)
            last = torch.full((num_nodes,), -1, dtype=torch.int64, device=dev).scatter_reduce_(0, row[pos], pos, "amax")
            has = last >= 0
            loop_w[has] = w[last[has]]
        nodes = torch.arange(num_nodes, device=dev)
        row, col, w = torch.cat([row[~loop], nodes]), torch.cat([col[~loop], nodes]), torch.cat([w[~loop], loop_w])
        deg = torch.zeros(num_nodes, dtype=torch.float32, device=dev).scatter_add_(0, col, w)   # at the target
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        norm = dis[row] * w * dis[col]
        self.n = num_nodes
        self.fwd = self._csr(col, row, norm, num_nodes)   # rows = targets, columns = sources
        self.bwd = self._csr(row, col, norm, num_nodes)   # transposed

    @staticmethod
    def _csr(rows, cols, vals, n):
        order = torch.sort(rows, stable=True).indices
        ptr_ = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
        ptr_[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
        return ptr_, cols[order].to(torch.int32).contiguous(), vals[order].contiguous()


_SPMM_MAX_D = 128   # gg_spmm_csr_f32: channels per call


def _spmm(csr, x: torch.Tensor) -> torch.Tensor:
    row_ptr, col, val = csr
    if int(row_ptr.numel()) - 1 != int(x.shape[0]):
        raise ValueError(f"adjacency has {int(row_ptr.numel()) - 1} nodes, features have {int(x.shape[0])} rows")
    lib = _lib.load()
    d = int(x.shape[1])
    if d <= _SPMM_MAX_D:
        x = x.contiguous()
        y = torch.empty_like(x)
        check(lib.gg_spmm_csr_f32(ptr(row_ptr), ptr(col), ptr(val), int(x.shape[0]), ptr(x), d, ptr(y), _stream()),
              "gg_spmm_csr_f32")
        return y
    # the kernel keeps one row of <= 128 channels per warp: wider layers go through it in column tiles
    parts = []
    for c0 in range(0, d, _SPMM_MAX_D):
        xt = x[:, c0:c0 + _SPMM_MAX_D].contiguous()
        yt = torch.empty_like(xt)
        check(lib.gg_spmm_csr_f32(ptr(row_ptr), ptr(col), ptr(val), int(x.shape[0]), ptr(xt), int(xt.shape[1]), ptr(yt),
                                  _stream()), "gg_spmm_csr_f32")
        parts.append(yt)
    return torch.cat(parts, dim=1)


class _Aggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, adj: NormAdj):
        ctx.adj = adj
        return _spmm(adj.fwd, x)

    @staticmethod
    def backward(ctx, grad):
        return _spmm(ctx.adj.bwd, grad), None


class GCNConv(torch.nn.Module):
    """x -> A_norm (x W) + b."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        torch.nn.init.xavier_uniform_(self.lin.weight)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, adj: NormAdj):
        return _Aggregate.apply(self.lin(x), adj) + self.bias


class GCN(torch.nn.Module):
    """gnn_models.py:58-100, same constructor arguments; ``forward(x, edges_config)`` takes the reference's dictionary
    of ``edge_index_<T>`` / ``edge_weight_<T>`` tensors (normalised adjacencies are cached per edge set)."""

    def __init__(self, num_features: int, layers_config: List[str], activation: str = "ReLU"):
        super().__init__()
        self.layers_config = interpret_layers_config(list(layers_config))
        chans = [num_features] + [c["channels"] for c in self.layers_config]
        self.convs = torch.nn.ModuleList(GCNConv(a, b) for a, b in zip(chans[:-1], chans[1:]))
        self.act = getattr(torch.nn, activation)()
        self._adj: Dict[str, NormAdj] = {}

    def adjacency(self, edges_config, cfg, n):
        """Normalised adjacency of the layer's edge set, cached on the IDENTITY of the tensors it was built from: the
        reference flow keeps the trained encoder and then runs it on a rebuilt, larger graph (kmer_ssgnn.py:127-146),
        and mini-batches hand the same encoder a different edges_config per call."""
        name = cfg["edge_index"]
        ei, w = edges_config[name], edges_config.get(cfg["edge_weight"])
        key = (ei.data_ptr(), tuple(ei.shape), ei._version, None if w is None else (w.data_ptr(), w._version), int(n))
        hit = self._adj.get(name)
        if hit is None or hit[0] != key:
            hit = (key, NormAdj(ei, w, n))
            self._adj[name] = hit
        return hit[1]

    def invalidate(self):
        """Drop the cached adjacencies (after editing an edge tensor in place)."""
        self._adj.clear()

    def forward(self, x, edges_config):
        for i, (conv, cfg) in enumerate(zip(self.convs, self.layers_config)):
            if i:
                x = self.act(x)
            x = conv(x, self.adjacency(edges_config, cfg, x.shape[0]))
        return x


# --------------------------------------------------------------------------- contrastive task
class SamplingTask:
    """Full-batch contrastive training of the encoder (sampling_task.py:236-382): every ``resample_every_num_epochs``
    epochs positive pairs are drawn from biased random walks on the De Bruijn graph (weight
    ``edge_weight_loss_weight``) and from every KF edge set in proportion to its weights
    (``kmer_frequency_loss_weight``), negatives uniformly (``negative_sampling_loss_weight``); the loss is the mean
    -logsigmoid of +/- the embedding dot products; Adam."""

    def __init__(self, graph, encoder: GCN, edges_config: Dict[str, torch.Tensor], edge_weight_loss_weight=1.0,
                 kmer_frequency_loss_weight=1.0, negative_sampling_loss_weight=1.0, proportion_nodes_to_sample=1.0,
                 proportion_negative_to_positive_samples=1.0, num_walks=5, walk_length=2, window_size=3, rw_p=1.0,
                 rw_q=1.0, lr=1e-2, resample_every_num_epochs=1, seed=0):
        _require_cuda()
        self.graph, self.encoder, self.edges_config = graph, encoder.cuda(), {k: v.cuda() for k, v in edges_config.items()}
        self.x = (graph.x if hasattr(graph, "x") else graph["x"]).cuda().float()
        self.w_db, self.w_kf, self.w_neg = edge_weight_loss_weight, kmer_frequency_loss_weight, negative_sampling_loss_weight
        self.prop_nodes, self.prop_neg = proportion_nodes_to_sample, proportion_negative_to_positive_samples
        self.num_walks, self.walk_length, self.window_size, self.rw_p, self.rw_q = num_walks, walk_length, window_size, rw_p, rw_q
        self.resample = resample_every_num_epochs
        self.opt = torch.optim.Adam(self.encoder.parameters(), lr=lr)
        self.gen = torch.Generator(device="cuda")
        self.gen.manual_seed(seed)
        self.seed, self.epoch = seed, 0
        self.kf_sets = sorted(k[len("edge_index_"):] for k in self.edges_config if k.startswith("edge_index_KF"))
        self.pairs: List[tuple] = []

    def sample_pairs(self):
        n_sample = int(self.edges_config["edge_index_DB"].shape[1] * self.prop_nodes)
        self.pairs, total = [], 0
        if self.w_db != 0:
            pos = sample_biased_random_walks(
                {"edge_index": self.edges_config["edge_index_DB"], "weight": self.edges_config["edge_weight_DB"],
                 "num_nodes": int(self.x.shape[0])}, n_sample, num_walks=self.num_walks, walk_length=self.walk_length,
                window_size=self.window_size, p=self.rw_p, q=self.rw_q, seed=self.seed * 1_000_003 + self.epoch)
            if pos.numel():
                self.pairs.append((pos, self.w_db, +1.0))
                n_sample = total = int(pos.shape[1])
        if self.w_kf != 0:
            for name in self.kf_sets:
                ei, w = self.edges_config["edge_index_" + name], self.edges_config["edge_weight_" + name]
                m = n_sample // max(len(self.kf_sets), 1)
                if ei.shape[1] == 0 or m == 0:
                    continue
                # np.random.choice(E, size=m, p=w / w.sum()) by inverse CDF (torch.multinomial stops at 2^24 categories)
                cdf = torch.cumsum(w.double(), 0)
                u = torch.rand(m, dtype=torch.float64, device="cuda", generator=self.gen) * cdf[-1]
                idx = torch.searchsorted(cdf, u, right=True).clamp_(max=int(w.numel()) - 1)
                self.pairs.append((ei[:, idx], self.w_kf, +1.0))
                total += m
        if self.w_neg != 0 and total:
            m = int(total * self.prop_neg)
            n = int(self.x.shape[0])
            a = torch.randint(0, n, (m,), device="cuda", generator=self.gen)
            b = (a + torch.randint(1, n, (m,), device="cuda", generator=self.gen)) % n   # uniform over pairs a != b
            self.pairs.append((torch.stack([a, b]), self.w_neg, -1.0))

    def loss(self, z):
        total = z.new_zeros(())
        for pairs, weight, sign in self.pairs:
            total = total - F.logsigmoid(sign * (z[pairs[0]] * z[pairs[1]]).sum(-1)).mean() * weight
        return total

    def fit(self, epochs: int) -> List[float]:
        history = []
        self.encoder.train()
        for _ in range(epochs):
            if self.epoch % self.resample == 0:
                self.sample_pairs()
            self.opt.zero_grad(set_to_none=True)
            loss = self.loss(self.encoder(self.x, self.edges_config))
            loss.backward()
            self.opt.step()
            history.append(float(loss.item()))
            self.epoch += 1
        return history

    @torch.no_grad()
    def inference(self) -> torch.Tensor:
        self.encoder.eval()
        return self.encoder(self.x, self.edges_config)


def edges_config_from(built) -> Dict[str, torch.Tensor]:
    """The reference's ``edges_config`` dictionary (sampling_task.py:118-161) from a ``core.BuiltGraph``."""
    cfg = {"edge_index_DB": built.db.edge_index, "edge_weight_DB": built.db.weight}
    for i, kf in enumerate(built.kf):
        cfg[f"edge_index_KF{i}"], cfg[f"edge_weight_KF{i}"] = kf.edge_index, kf.weight
    return cfg


# --------------------------------------------------------------------------- downstream: edit-distance approximation
def hyperbolic_distance(u, v, epsilon=1e-7):
    """distances.py:20-28."""
    sqdist = torch.sum((u - v) ** 2, dim=-1)
    x = 1 + 2 * sqdist / ((1 - torch.sum(u**2, dim=-1)) * (1 - torch.sum(v**2, dim=-1))) + epsilon
    return torch.log(x + torch.sqrt(x**2 - 1))


class EditDistanceMLP(torch.nn.Module):
    """Single-k flavour of edit_distance_models/mlp.py: frozen k-mer embedding (one extra all-zero row for unknown /
    N-holding k-mers, kmer_ssgnn.py:316-330) -> flatten -> dense layers -> point in the Poincare ball -> hyperbolic
    distance between the two sequences of a pair, trained with MSE against the (scaled) edit distance."""

    def __init__(self, kmer_embeddings: torch.Tensor, n_kmers_per_sequence: int, layers=(128, 32), activation="ReLU"):
        super().__init__()
        weights = torch.cat([kmer_embeddings, kmer_embeddings.new_zeros(1, kmer_embeddings.shape[1])])
        self.embedding = torch.nn.Embedding.from_pretrained(weights, freeze=True)
        dims = [n_kmers_per_sequence * kmer_embeddings.shape[1], *layers]
        mods = []
        for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
            if i:
                mods.append(getattr(torch.nn, activation)())
            mods.append(torch.nn.Linear(a, b))
        self.net = torch.nn.Sequential(*mods)

    def encode(self, kmers_index):
        e = self.net(self.embedding(kmers_index).flatten(1))
        norm = e.norm(dim=-1, keepdim=True)
        return e / (1 + norm)     # inside the unit ball

    def forward(self, a_index, b_index):
        return hyperbolic_distance(self.encode(a_index), self.encode(b_index))


def fit_edit_distance(model: EditDistanceMLP, a_index, b_index, distances, epochs=200, lr=1e-3, seed=0):
    """Full-batch MSE training; returns the loss history."""
    torch.manual_seed(seed)
    for m in model.net:
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            torch.nn.init.zeros_(m.bias)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr)
    history = []
    for _ in range(epochs):
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model(a_index, b_index), distances)
        loss.backward()
        opt.step()
        history.append(float(loss.item()))
    return history
