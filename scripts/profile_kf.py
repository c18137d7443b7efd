"""Run only the KF emit kernels (ncu target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s = int(sys.argv[2]) if len(sys.argv) > 2 else 5
impl = sys.argv[3] if len(sys.argv) > 3 else "auto"
reads = synthetic_reads(200_000, 150, 0)
built = core.build_graph(reads, k, str(s), with_kf=False)
kc = built.kf_counts[0]
for _ in range(3):
    recs, n, rowcnt, params, scored = core.kf_emit(kc.counts, kc.sqnorm, 0.8, kc.m ** 2, impl=impl, rec_capacity=1 << 25)
torch.cuda.synchronize()
print("candidates", n)
