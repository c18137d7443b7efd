"""KF narrow-row paths at k=8 (development aid): python scripts/kfprof.py [s] [impl]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

s = int(sys.argv[1]) if len(sys.argv) > 1 else 2
impl = sys.argv[2] if len(sys.argv) > 2 else "dedup"
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
reads = synthetic_reads(1_000_000 if k >= 8 else 200_000, 150, seed=0)
built = core.build_graph(reads, k, str(s), with_kf=False)
kc = built.kf_counts[0]
for it in range(3):
    timer = core.StageTimer(); core.set_timer(timer)
    kf = core.kf_edges(kc.counts, kc.sqnorm, 0.8, 0.01, sq_max=kc.m ** 2, impl=impl)
    core.set_timer(None)
    t = {k_: round(v[0], 4) for k_, v in timer.summary().items()}
print(impl, "E =", kf.edge_index.shape[1], t, "sum", round(sum(t.values()), 3))
