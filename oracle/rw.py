"""CPU restatement of the reference's biased random-walk sampler (TEST INFRASTRUCTURE ONLY -- never imported by the
product path; see oracle/__init__.py).

Reference: src/representations/gnn_common/gnn_utils.py:514-600 (sample_biased_random_walks).  The reference draws from
Python's ``random``, so what can be pinned is the LAW of its output: ``walk_law`` enumerates every walk from a start
node with its exact probability, ``pairs_per_walk_law`` gives the exact distribution of the number of pairs a walk of a
given length emits, and ``pairs_of_walk`` restates the window rule for given half-widths.  ``sample`` is the line by
line restatement (networkx-free) used to cross-check those closed forms.
"""
import random
from collections import defaultdict
from fractions import Fraction


def successors(edge_index, weight, n):
    """Successor lists in edge order -- what to_networkx(data) + G.neighbors(u) yield (gnn_utils.py:537, 554)."""
    nbrs = [[] for _ in range(n)]
    for (u, v), w in zip(zip(*[list(map(int, r)) for r in edge_index]), weight):
        nbrs[u].append((v, float(w)))
    return nbrs


def step_weights(nbrs, cur, prev, p, q):
    """Unnormalised transition weights of one step (gnn_utils.py:559-583)."""
    out = []
    for v, w in nbrs[cur]:
        if prev is None:
            out.append(w)                                   # :560-561 first step: plain edge weights
        elif v == prev:
            out.append(w / p)                               # :566-569 return
        elif any(x == prev for x, _ in nbrs[v]):
            out.append(w)                                   # :570-573 G.has_edge(neighbor, previous_node)
        else:
            out.append(w / q)                               # :574-577 in-out
    return out


def walk_law(nbrs, start, walk_length, p=1.0, q=1.0):
    """{walk tuple: probability} over all walks from ``start`` (a node without successors ends the walk, :585-586)."""
    law = defaultdict(float)

    def rec(walk, prev, prob):
        cur = walk[-1]
        if len(walk) == walk_length + 1 or not nbrs[cur]:
            law[tuple(walk)] += prob
            return
        ws = step_weights(nbrs, cur, prev, p, q)
        total = sum(ws)
        for (v, _), w in zip(nbrs[cur], ws):
            if w > 0:
                rec(walk + [v], cur, prob * w / total)

    rec([start], None, 1.0)
    return dict(law)


def pairs_of_walk(walk, shrinks):
    """(walk[i], walk[j]) for j in [i - shrink_i, i + shrink_i], j != i (gnn_utils.py:591-596)."""
    out = []
    for i in range(len(walk)):
        for j in range(i - shrinks[i], i + shrinks[i] + 1):
            if 0 <= j < len(walk) and j != i:
                out.append((walk[i], walk[j]))
    return out


def pairs_per_walk_law(n_nodes_in_walk, window_size):
    """Exact distribution {number of pairs: probability} for a walk with that many nodes, shrink_i ~ U{1..window}."""
    law = {0: Fraction(1)}
    for i in range(n_nodes_in_walk):
        nxt = defaultdict(Fraction)
        for s in range(1, window_size + 1):
            c = min(i, s) + min(n_nodes_in_walk - 1 - i, s)
            for tot, pr in law.items():
                nxt[tot + c] += pr / window_size
        law = dict(nxt)
    return {k: float(v) for k, v in law.items()}


def sample(edge_index, weight, n, num_nodes_to_sample, num_walks=5, walk_length=2, window_size=3, p=1.0, q=1.0,
           rng=random):
    """Line-by-line restatement of gnn_utils.py:537-598; returns (pairs, walks)."""
    nbrs = successors(edge_index, weight, n)
    all_nodes = list(range(n))
    nodes = rng.sample(all_nodes, num_nodes_to_sample) if num_nodes_to_sample < n else all_nodes
    rng.shuffle(nodes)
    pairs, walks = [], []
    for _ in range(num_walks):
        for node in nodes:
            walk, prev = [node], None
            for _ in range(walk_length):
                cur = walk[-1]
                if not nbrs[cur]:
                    break
                ws = step_weights(nbrs, cur, prev, p, q)
                if prev is not None:
                    tot = sum(ws)
                    ws = [x / tot for x in ws]
                walk.append(rng.choices([v for v, _ in nbrs[cur]], weights=ws, k=1)[0])
                prev = cur
            shrinks = [rng.randint(1, window_size) for _ in walk]
            pairs.extend(pairs_of_walk(walk, shrinks))
            walks.append(walk)
    return pairs, walks
