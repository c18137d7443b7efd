"""Where does the host time of one step go? (development aid)"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

reads = torch.from_numpy(synthetic_reads(1_000_000, 150, 0)).pin_memory()
stream = core.upload_reads(reads)
kw = dict(k=8, small_k="2,5", edges_threshold=0.8, edges_keep_top_k=0.01)
for _ in range(3):
    core.build_graph(stream, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    core.build_graph(stream, **kw)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    core.build_graph(stream, **kw)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(35)
