"""Row f3: the GPU biased random-walk sampler against the law of the reference's sampler
(gnn_utils.py:514-600) as stated by oracle/rw.py, which tests/test_oracle.py pins to statistics of the real
reference.  The reference draws from Python's ``random``: parity is distributional, plus exact structural checks of
the pair list against the walks the kernel reports."""
import itertools
import os
import types

import numpy as np
import pytest
import torch

from oracle import rw
from tests._util import GOLDEN

pytestmark = pytest.mark.gpu


def _graph():
    z = np.load(os.path.join(GOLDEN, "rw_reference.npz"))
    n = int(z["num_nodes"])
    data = types.SimpleNamespace(edge_index=torch.from_numpy(z["edge_index"]), weight=torch.from_numpy(z["weight"]),
                                 num_nodes=n)
    return z, n, data


def test_walk_law_and_window_law():
    import genomic_gnn_b200 as gg

    z, n, data = _graph()
    nbrs = rw.successors(z["edge_index"], z["weight"], n)
    p, q = float(z["p"]), float(z["q"])
    rounds = 4000
    pairs, walks, _ = gg.sample_biased_random_walks(data, n, num_walks=rounds, walk_length=2, window_size=1, p=p, q=q,
                                                    seed=11, return_walks=True)
    walks = walks.numpy()
    assert walks.shape == (rounds * n, 3) and pairs.shape == (2, 4 * rounds * n) and pairs.dtype == torch.long
    # every round visits every node once, in the same shuffled order
    assert np.array_equal(np.sort(walks[:n, 0]), np.arange(n))
    assert np.array_equal(walks[:, 0].reshape(rounds, n), np.tile(walks[:n, 0], (rounds, 1)))
    keys, counts = np.unique(walks, axis=0, return_counts=True)
    got = {tuple(int(x) for x in k): int(c) for k, c in zip(keys, counts)}
    chi2, dof = 0.0, 0
    for a in range(n):
        law = rw.walk_law(nbrs, a, 2, p=p, q=q)
        assert set(k for k in got if k[0] == a) <= set(law)
        for walk, pr in law.items():
            exp = rounds * pr
            assert abs(got.get(walk, 0) - exp) <= 5.0 * np.sqrt(exp * (1 - pr)) + 1.0, (walk, got.get(walk, 0), exp)
            if exp >= 5:
                chi2 += (got.get(walk, 0) - exp) ** 2 / exp
                dof += 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
    # window_size = 1: the pair list is exactly (a,b),(b,a),(b,c),(c,b) per walk, walks in order
    want = np.stack([walks[:, [0, 1, 1, 2]].reshape(-1), walks[:, [1, 0, 2, 1]].reshape(-1)])
    assert np.array_equal(pairs.numpy(), want)

    # window law: pairs per walk, and every walk's pairs are a legal window pattern
    pairs, walks, counts = gg.sample_biased_random_walks(data, n, num_walks=1500, walk_length=4, window_size=3, seed=5,
                                                         return_walks=True)
    walks, pl, counts = walks.numpy(), [tuple(x) for x in pairs.t().tolist()], counts.numpy()
    assert counts.sum() == len(pl)
    law = rw.pairs_per_walk_law(5, 3)
    patterns = {}
    for shrinks in itertools.product((1, 2, 3), repeat=5):
        patterns.setdefault(tuple(rw.pairs_of_walk(list(range(5)), shrinks)), shrinks)
    t, hist = 0, {}
    for w, c in zip(walks, counts):
        seg = pl[t:t + c]
        t += c
        # some choice of half-widths reproduces exactly this walk's pairs (positions stand in for nodes)
        assert any(len(pat) == c and all((int(w[i]), int(w[j])) == sg for (i, j), sg in zip(pat, seg))
                   for pat in patterns), (w, seg)
        hist[int(c)] = hist.get(int(c), 0) + 1
    total = len(walks)
    assert set(hist) <= set(law)
    for k, pr in law.items():
        exp = total * pr
        assert abs(hist.get(k, 0) - exp) <= 5.0 * np.sqrt(exp * (1 - pr)) + 1.0, (k, hist.get(k, 0), exp)


def test_reproducible_subsample_dead_ends_and_device():
    import genomic_gnn_b200 as gg

    z, n, data = _graph()
    a = gg.sample_biased_random_walks(data, 5, num_walks=3, walk_length=3, window_size=2, p=2.0, q=0.5, seed=1)
    b = gg.sample_biased_random_walks(data, 5, num_walks=3, walk_length=3, window_size=2, p=2.0, q=0.5, seed=1)
    c = gg.sample_biased_random_walks(data, 5, num_walks=3, walk_length=3, window_size=2, p=2.0, q=0.5, seed=2)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert not a.is_cuda                                    # comes back where the graph lives
    assert len(set(a[0].tolist())) <= 16 and a.shape[0] == 2
    # a path graph 0 -> 1 -> 2 with a dead end: walks stop there; an edgeless graph yields the reference's empty tensor
    path = types.SimpleNamespace(edge_index=torch.tensor([[0, 1], [1, 2]]), weight=torch.tensor([1.0, 1.0]), num_nodes=3)
    pairs, walks, counts = gg.sample_biased_random_walks(path, 3, num_walks=1, walk_length=5, window_size=1, seed=0,
                                                         return_walks=True)
    assert sorted(counts.tolist()) == [0, 2, 4]
    by_start = {int(w[0]): [int(x) for x in w if x >= 0] for w in walks}
    assert by_start == {0: [0, 1, 2], 1: [1, 2], 2: [2]}
    assert pairs.shape[1] == 4 + 2 + 0
    none = types.SimpleNamespace(edge_index=torch.zeros((2, 0), dtype=torch.long), weight=torch.zeros(0), num_nodes=4)
    out = gg.sample_biased_random_walks(none, 4)
    assert out.numel() == 0 and out.dtype == torch.long
    with pytest.raises(gg.GGError):
        gg.sample_biased_random_walks(data, 4, walk_length=64)


def test_sampler_on_a_built_graph():
    """End to end on the graph the builders hand out: every emitted pair lies within window_size steps of a real walk."""
    import genomic_gnn_b200 as gg
    from genomic_gnn_b200 import core
    from oracle import fast

    built = core.build_graph(fast.synthetic_reads(3000, 150, seed=4), 6, "2", with_kf=False)
    data = types.SimpleNamespace(edge_index=built.db.edge_index, weight=built.db.weight, num_nodes=built.db.num_nodes)
    pairs, walks, _ = gg.sample_biased_random_walks(data, 1000, num_walks=5, walk_length=2, window_size=3, p=1.0, q=1.0,
                                                    seed=9, return_walks=True)
    assert pairs.is_cuda and walks.shape == (5000, 3)
    edges = set(map(tuple, built.db.edge_index.t().tolist()))
    w = walks.cpu().numpy()
    assert all((int(a), int(b)) in edges for a, b in zip(w[:, 0], w[:, 1]))
    assert all((int(a), int(b)) in edges for a, b in zip(w[:, 1], w[:, 2]))
    assert 4 * 5000 <= pairs.shape[1] <= 6 * 5000          # 3-node walks: 4 pairs at shrink 1, 6 when every window covers the walk


def test_mini_batch_wrapper_forwards_like_the_reference():
    """gnn_tasks_miniBatch/utils.py:43-45 forwards positionally: (walk_length, window_size, p, q) land in
    (num_walks, walk_length, window_size, p)."""
    import genomic_gnn_b200 as gg

    z, n, data = _graph()
    hetero = {("node", "DB", "node"): types.SimpleNamespace(edge_index=data.edge_index, edge_attr=data.weight),
              "node": types.SimpleNamespace(x=torch.zeros((n, 4)))}
    a = gg.sample_biased_random_walks_miniBatch(hetero, n, num_walks=7, walk_length=3, window_size=2, p=1.0, q=0.25, seed=4)
    b = gg.sample_biased_random_walks(data, n, num_walks=3, walk_length=2, window_size=1, p=0.25, q=1.0, seed=4)
    assert torch.equal(a, b) and a.shape == (2, 4 * 3 * n)
