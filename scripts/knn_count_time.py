"""Pass-1 timing of the per-row kNN at BASELINE config 4 size (development aid): python scripts/knn_count_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from genomic_gnn_b200 import core, _lib
from genomic_gnn_b200._lib import ptr, check
from genomic_gnn_b200.synth import synthetic_reads

built = core.build_graph(synthetic_reads(200_000, 150, seed=0), 8, "5", with_kf=False)
kc = built.kf_counts[0]
n, pitch = kc.counts.shape
lib = _lib.load()
sqnorm = kc.sqnorm.contiguous()
sq_values = tuple(int(v) for v in torch.unique(sqnorm).cpu().tolist())
sqid, tab, rank_q, n_ranks = core.knn_tables(sq_values, "IP")
dev = kc.counts.device
sqid_d, tab_d = torch.from_numpy(sqid).to(dev), torch.from_numpy(tab.view(np.int16)).to(dev)
hist = torch.empty(n * n_ranks, dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    check(lib.gg_knn_count(ptr(kc.counts), pitch, n, 0, n, ptr(sqnorm), ptr(sqid_d), int(sq_values[-1]), ptr(tab_d),
                           int(tab.shape[0]), int(tab.shape[1]) - 1, n_ranks, -1, ptr(hist), st), "count")
for _ in range(2): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): run()
b.record(); torch.cuda.synchronize()
print(f"N {n} D {pitch} n_ranks {n_ranks} n_sq {tab.shape[0]}: pass 1 (tcgen05) {a.elapsed_time(b)/5:.3f} ms")
