import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads
reads = torch.from_numpy(synthetic_reads(1_000_000, 150, 0)).pin_memory()
kw = dict(k=8, small_k="2,5", edges_threshold=0.8, edges_keep_top_k=0.01)
last = None
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = core.build_graph_to_host(reads, **kw)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    ms = torch.cuda.memory_stats()
    print(f"step {i}: {dt:.2f} ms  reserved={ms['reserved_bytes.all.current']/1e9:.2f} GB  cudaMalloc calls={ms['num_device_alloc']}  retries={ms['num_alloc_retries']}", flush=True)
    last = g
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3):
    last = core.build_graph_to_host(reads, **kw)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
