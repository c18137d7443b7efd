"""Dense restatements for row f1 (TEST INFRASTRUCTURE ONLY).

``gcn_conv_dense`` restates torch_geometric 2.3.1 ``GCNConv`` (what gnn_common/gnn_models.py:58-100 stacks) with plain
dense float64 algebra: gcn_norm(add_self_loops=True, improved=False) -- remaining self loops of weight 1, degree summed
at the TARGET of every edge, norm = deg^-1/2[source] * w * deg^-1/2[target] -- then out = A_norm (x W^T) + b with
messages flowing source -> target.  torch_geometric is not installed in this image, so this follows the published
source of that version; the reference holds no test that pins it (parity unpinned).
``edit_distance`` is the textbook Levenshtein DP the downstream task regresses on.
"""
import numpy as np


def gcn_norm_dense(edge_index, edge_weight, n):
    row, col = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    w = np.ones(len(row)) if edge_weight is None else np.asarray(edge_weight, dtype=np.float64)
    loop_w = np.ones(n)
    keep = row != col
    for r, x in zip(row[~keep], w[~keep]):
        loop_w[r] = x                                   # an existing self loop keeps its weight
    row = np.concatenate([row[keep], np.arange(n)])
    col = np.concatenate([col[keep], np.arange(n)])
    w = np.concatenate([w[keep], loop_w])
    deg = np.zeros(n)
    np.add.at(deg, col, w)
    with np.errstate(divide="ignore"):
        dis = deg ** -0.5
    dis[np.isinf(dis)] = 0.0
    a = np.zeros((n, n))
    np.add.at(a, (col, row), dis[row] * w * dis[col])   # a[target, source]
    return a


def gcn_conv_dense(x, weight, bias, edge_index, edge_weight):
    a = gcn_norm_dense(edge_index, edge_weight, x.shape[0])
    return a @ (np.asarray(x, dtype=np.float64) @ np.asarray(weight, dtype=np.float64).T) + np.asarray(bias, dtype=np.float64)


def edit_distance(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]
