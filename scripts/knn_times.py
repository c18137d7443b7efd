"""Row f4b timings on the BASELINE config 4 graph (development aid): exact per-row kNN KF edges, k = int(0.01 N).
python scripts/knn_times.py [impl ...]   (default: auto sparse tcgen05 mma for the wide set, auto for the narrow one)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

built = core.build_graph(synthetic_reads(1_000_000, 150, seed=0), 8, "2,3,4,5", with_kf=False)
n = built.db.num_nodes
k_nn = int(0.01 * n)
impls = sys.argv[1:] or ["auto", "sparse", "tcgen05", "mma"]
for kc in built.kf_counts:
    d = int(kc.counts.shape[1])
    for impl in impls:
        if (impl == "sparse" and d < 64) or (impl == "tcgen05" and d < 128):
            continue
        for distance in ("IP", "L2"):
            core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, impl=impl)      # warm-up
            timer = core.StageTimer()
            core.set_timer(timer)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, distance, impl=impl)
            b.record()
            torch.cuda.synchronize()
            core.set_timer(None)
            stages = {k: round(sum(v), 3) for k, v in timer.summary().items()}
            print(f"N {n} D {d} {impl:8s} {distance} k {k_nn}: {a.elapsed_time(b):7.2f} ms, edges {res.edge_index.shape[1]}, "
                  f"pairs/s {n * (n - 1) / a.elapsed_time(b) / 1e-3:.3e}  stages {stages}", flush=True)
