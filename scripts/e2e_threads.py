"""Host-to-host build: host threads for the feature expansion (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads
reads = torch.from_numpy(synthetic_reads(1_000_000, 150, 0)).pin_memory()
kw = dict(k=8, small_k="2,5", edges_threshold=0.8, edges_keep_top_k=0.01)
for nt in (16, 8, 4, 12, 16, 8):
    core._HOST_THREADS = nt
    ts, last = [], None
    for i in range(25):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = core.build_graph_to_host(reads, **kw)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        last = g
    ts = sorted(ts[5:])
    print(f"threads {nt:2d}: median {ts[len(ts)//2]:.2f}  min {ts[0]:.2f}  p90 {ts[int(len(ts)*0.9)]:.2f}", flush=True)
core._HOST_THREADS = 16
ts = []
for i in range(25):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = core.build_graph_to_host(reads, host_expand=False, **kw)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
ts = sorted(ts[5:]); print(f"no host expansion: median {ts[len(ts)//2]:.2f}  min {ts[0]:.2f}  p90 {ts[int(len(ts)*0.9)]:.2f}")
