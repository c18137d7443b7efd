"""Freeze statistics of the REAL reference sampler (gnn_utils.py:514-600) for tests/golden/rw_reference.npz.

Runs only in the dev container (needs /root/reference).  torch_geometric is absent there; the one function the sampler
takes from it, ``to_networkx``, is stubbed with its documented behaviour (nodes 0..n-1, directed edges in edge order,
edge attribute ``weight``).  Two experiments, both on the De Bruijn graph of 40 seeded 60-bp reads at k=2:
  * walk law: walk_length=2, window_size=1 -> every 3-node walk emits (a,b),(b,a),(b,c),(c,b), so the pair stream
    decodes back into walks; p=0.5, q=2.0 exercise all three node2vec branches;
  * window law: one walk per call (walk_length=4, window_size=3) -> number of pairs per walk.
"""
import os
import random
import sys
import types
from collections import Counter

import networkx as nx
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import faithful, ref_import  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rw_reference.npz")


def to_networkx(data, edge_attrs=None, to_undirected=False):
    g = nx.DiGraph()
    g.add_nodes_from(range(int(data.num_nodes)))
    for (u, v), w in zip(data.edge_index.t().tolist(), data.weight.tolist()):
        g.add_edge(u, v, weight=w)
    return g


def main():
    ref_import.load()
    import src.representations.gnn_common.gnn_utils as gu

    gu.to_networkx = to_networkx
    rng = np.random.default_rng(7)
    reads = ["".join("ACGT"[c] for c in rng.integers(0, 4, 60)) for _ in range(40)]
    g = faithful.build_reference_graph(reads, 2, "1")
    ei, w = np.asarray(g["edge_index"]), np.asarray(g["weight"], dtype=np.float32)
    n = int(g["num_nodes"])
    data = types.SimpleNamespace(edge_index=torch.from_numpy(ei), weight=torch.from_numpy(w), num_nodes=n)

    random.seed(12345)
    walks = Counter()
    for _ in range(400):
        pairs = gu.sample_biased_random_walks(data, n, num_walks=5, walk_length=2, window_size=1, p=0.5, q=2.0)
        pl = pairs.t().tolist()
        assert len(pl) % 4 == 0          # every node of this graph has successors: all walks have 3 nodes
        for t in range(0, len(pl), 4):
            (a, b), (b2, a2), (b3, c), (c2, b4) = pl[t:t + 4]
            assert a == a2 and b == b2 == b3 == b4 and c == c2
            walks[(a, b, c)] += 1
    keys = sorted(walks)
    per_walk = Counter()
    for _ in range(20000):
        pairs = gu.sample_biased_random_walks(data, 1, num_walks=1, walk_length=4, window_size=3, p=1.0, q=1.0)
        per_walk[int(pairs.shape[1]) if pairs.dim() == 2 else 0] += 1
    np.savez_compressed(GOLDEN, edge_index=ei, weight=w, num_nodes=n, p=0.5, q=2.0,
                        walk_keys=np.array(keys, dtype=np.int64), walk_counts=np.array([walks[k] for k in keys]),
                        pairs_per_walk=np.array(sorted(per_walk.items()), dtype=np.int64))
    print("wrote", GOLDEN, len(keys), "distinct walks,", sum(walks.values()), "walks;", dict(per_walk))


if __name__ == "__main__":
    main()
