"""ncu target: one exact-kNN KF build (IP) per small_k on the BASELINE config 4 nodes (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

built = core.build_graph(synthetic_reads(200_000, 150, seed=0), 8, "2,5", with_kf=False)
n = built.db.num_nodes
for kc in (built.kf_counts[::-1] if len(sys.argv) < 2 else built.kf_counts):
    res = core.kf_knn_edges(kc.counts, kc.sqnorm, int(0.01 * n), "IP")
    torch.cuda.synchronize()
    print(n, kc.counts.shape, res.edge_index.shape)
