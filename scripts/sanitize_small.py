"""Small end-to-end run of the newer kernels for compute-sanitizer (development aid):
compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genomic_gnn_b200 import core, sample_biased_random_walks
from genomic_gnn_b200.synth import synthetic_reads

reads = synthetic_reads(3000, 150, seed=1)
reads[::37, 40] = ord("N")
stream = core.upload_reads(reads)
for k in (3, 6, 7, 8):
    a = core.count_kp1mers(stream, k, impl="part")
    b = core.count_kp1mers(stream, k, impl="l2")
    assert torch.equal(a.hist, b.hist) and torch.equal(a.first_pos, b.first_pos), k
ragged = ["ACGTNACGTTGCA" * 7, "ACG", "", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "ACGTACGTAGCTAGCTAGCATCGATCGATCGACTAGC" * 3]
for k in (3, 8):
    s2 = core.upload_reads(ragged * 50)
    a = core.count_kp1mers(s2, k, impl="part"); b = core.count_kp1mers(s2, k, impl="l2")
    assert torch.equal(a.hist, b.hist) and torch.equal(a.first_pos, b.first_pos) and a.n_bridges == b.n_bridges
built = core.build_graph(reads, 6, "2,4", edges_threshold=0.8, edges_keep_top_k=0.01)
kc = built.kf_counts[0]
for lo, hi in ((0, None), (1, 3)):
    x = core.kf_edges(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m ** 2, impl="dedup", iblock_begin=lo, iblock_end=hi)
    y = core.kf_edges(kc.counts, kc.sqnorm, 0.8, None, sq_max=kc.m ** 2, impl="dp4a", iblock_begin=lo, iblock_end=hi)
    assert torch.equal(x.edge_index, y.edge_index) and torch.equal(x.weight64, y.weight64)
x = core.kf_edges(kc.counts, kc.sqnorm, 0.5, 0.001, sq_max=kc.m ** 2, impl="dedup")
host = core.build_graph_to_host(torch.from_numpy(reads), 6, "2,4", edges_threshold=0.8, edges_keep_top_k=0.01)
assert torch.equal(host.x[1], built.x[1].cpu())
data = types.SimpleNamespace(edge_index=built.db.edge_index, weight=built.db.weight, num_nodes=built.db.num_nodes)
p = sample_biased_random_walks(data, 500, num_walks=3, walk_length=4, window_size=3, p=0.5, q=2.0, seed=3)
assert p.shape[0] == 2
torch.cuda.synchronize()
print("sanitize run ok")
