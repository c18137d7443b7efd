"""GPU idle gaps inside one build_graph step from a torch.profiler (CUPTI) kernel timeline (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

reads = synthetic_reads(1_000_000, 150, seed=0)
stream = core.upload_reads(reads)
kw = dict(k=8, small_k="2,5", edges_threshold=0.8, edges_keep_top_k=0.01)
for _ in range(3):
    core.build_graph(stream, **kw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    core.build_graph(stream, **kw)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print(f"{len(evs)} device activities, span {(evs[-1].time_range.end - t0)/1e3:.3f} ms, busy {busy/1e3:.3f} ms")
prev_end, prev_name = evs[0].time_range.end, evs[0].name
gaps = []
for e in evs[1:]:
    gap = e.time_range.start - prev_end
    if gap > 8:
        gaps.append((gap, (prev_end - t0) / 1e3, prev_name[:50], e.name[:50]))
    if e.time_range.end > prev_end:
        prev_end, prev_name = e.time_range.end, e.name
print("gaps > 8 us:", len(gaps), "total", sum(g[0] for g in gaps) / 1e3, "ms")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"{g[0]:7.1f} us at {g[1]:.3f} ms  after {g[2]}  -> before {g[3]}")
if len(sys.argv) > 1:   # full timeline: start (ms), duration (us), gap before (us), name
    with open(sys.argv[1], "w") as fh:
        prev = evs[0].time_range.start
        for e in evs:
            fh.write(f"{(e.time_range.start - t0) / 1e3:8.3f} {e.time_range.end - e.time_range.start:8.1f} {e.time_range.start - prev:7.1f}  {e.name[:90]}\n")
            prev = max(prev, e.time_range.end)
