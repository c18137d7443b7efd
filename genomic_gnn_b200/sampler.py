"""Row f3: positive pairs from biased random walks on the De Bruijn graph, on the GPU.

Drop-in for ``sample_biased_random_walks`` (reference gnn_common/gnn_utils.py:514-600), which the sampling tasks call
once per epoch (full batch, gnn_tasks/sampling_task.py:250-260) or once per batch (mini batch).  Same arguments, same
return value -- a ``torch.long`` tensor ``[2, P]`` of (walk[i], walk[j]) pairs -- but the walks run one per thread in
``gg_rw_pairs`` instead of a Python loop over networkx.  The reference draws from Python's ``random``; parity is
therefore distributional (tests/test_gpu_sampler.py checks the walk law against exact probabilities and the window
law against the walks), and ``seed`` makes a run reproducible.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .core import _require_cuda, _stream


def _field(data, name):
    if hasattr(data, name):
        return getattr(data, name)
    return data[name]


def sample_biased_random_walks(data, num_nodes_to_sample: int, num_walks: int = 5, walk_length: int = 2,
                               window_size: int = 3, p: float = 1.0, q: float = 1.0, seed=None,
                               return_walks: bool = False):
    """``data``: anything with ``edge_index`` (long [2, E]), ``weight`` (float [E]) and ``num_nodes`` -- the
    ``torch_graph`` that ``create_graphs`` builds.  Returns the pairs on ``edge_index``'s device (with
    ``return_walks`` also the walks, ``[n_walks, walk_length + 1]``, -1 padded, and the number of pairs of each walk)."""
    _require_cuda()
    lib = _lib.load()
    edge_index = _field(data, "edge_index")
    weight = _field(data, "weight")
    n = int(_field(data, "num_nodes"))
    home = edge_index.device
    dev = torch.device("cuda", torch.cuda.current_device())
    ei = edge_index.to(dev)
    w = weight.to(dev, torch.float32).contiguous()
    # CSR with the successors of a node in edge order (what to_networkx + G.neighbors yield)
    order = torch.sort(ei[0], stable=True).indices
    col = ei[1][order].contiguous()
    w = w[order].contiguous()
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(torch.bincount(ei[0], minlength=n), 0)
    gen = torch.Generator(device=dev)
    seed = int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed)
    gen.manual_seed(seed)
    # random.sample + random.shuffle (gnn_utils.py:540-545): a uniformly random ordered subset
    start = torch.randperm(n, generator=gen, device=dev)[: max(0, min(int(num_nodes_to_sample), n))].contiguous()
    n_start = int(start.numel())
    n_walks = n_start * int(num_walks)
    empty = torch.empty(0, dtype=torch.long, device=home)  # torch.tensor([]).t() in the reference
    if n_walks == 0:
        if return_walks:
            return empty, torch.empty((0, walk_length + 1), dtype=torch.long, device=home), empty.clone()
        return empty
    if ei.shape[1] == 0:  # no edges: every walk is its start node alone and emits nothing
        if return_walks:
            walks = torch.full((n_walks, walk_length + 1), -1, dtype=torch.long, device=dev)
            walks[:, 0] = start.repeat(int(num_walks))
            return empty, walks.to(home), torch.zeros(n_walks, dtype=torch.long, device=home)
        return empty
    counts = torch.empty(n_walks, dtype=torch.int64, device=dev)
    args = (ptr(row_ptr), ptr(col), ptr(w), n, ptr(start), n_start, int(num_walks), int(walk_length), int(window_size),
            float(p), float(q), C.c_uint64(seed & (2**64 - 1)))
    check(lib.gg_rw_pairs(*args, 0, ptr(counts), None, None, None, 0, None, _stream()), "gg_rw_pairs(count)")
    offsets = torch.cumsum(counts, 0) - counts
    total = int((offsets[-1] + counts[-1]).item())
    pairs = torch.empty((2, max(total, 1)), dtype=torch.int64, device=dev)
    walks = torch.empty((n_walks, walk_length + 1), dtype=torch.int64, device=dev) if return_walks else None
    check(lib.gg_rw_pairs(*args, 1, None, ptr(offsets), ptr(pairs[0]), ptr(pairs[1]), total, ptr(walks), _stream()),
          "gg_rw_pairs(write)")
    out = pairs[:, :total].to(home) if total else empty
    return (out, walks.to(home), counts.to(home)) if return_walks else out


def sample_biased_random_walks_miniBatch(data, num_nodes_to_sample: int, num_walks: int = 5, walk_length: int = 2,
                                         window_size: int = 3, p: float = 1.0, q: float = 1.0, seed=None):
    """Mini-batch flavour (reference gnn_tasks_miniBatch/utils.py:10-45): the walks run on the ("node", "DB", "node")
    edges of the sampled HeteroData batch.  The reference forwards its arguments POSITIONALLY as
    ``(graph, num_nodes_to_sample, walk_length, window_size, p, q)`` (utils.py:43-45), so the callee receives
    num_walks=walk_length, walk_length=window_size, window_size=p and p=q while ``num_walks`` and ``q`` of this
    wrapper are dropped; that behaviour is reproduced, not corrected."""
    store = data["node", "DB", "node"]
    graph = {"edge_index": store.edge_index if hasattr(store, "edge_index") else store["edge_index"],
             "weight": store.edge_attr if hasattr(store, "edge_attr") else store["edge_attr"],
             "num_nodes": int((data["node"].x if hasattr(data["node"], "x") else data["node"]["x"]).shape[0])}
    return sample_biased_random_walks(graph, num_nodes_to_sample, walk_length, window_size, int(p), q, seed=seed)
