"""K1 timings of the three counting paths (development aid): python scripts/count_times.py [R]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reads = synthetic_reads(R, 150, seed=0)
stream = core.upload_reads(reads)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for k in (5, 6, 7, 8, 9):
    ref = None
    for impl in ("l2", "smem", "part"):
        try:
            core.count_kp1mers(stream, k, impl=impl, first_occurrence=False)
        except ValueError:
            continue
        ts = []
        for _ in range(5):
            flush.fill_(1)
            timer = core.StageTimer(); core.set_timer(timer)
            cr = core.count_kp1mers(stream, k, impl=impl, first_occurrence=False)
            core.set_timer(None)
            ts.append(timer.summary()["k1_count"][0])
        if ref is None:
            ref = cr.hist.clone()
        ok = torch.equal(ref, cr.hist)
        ms = min(ts)
        print(f"k={k} {impl:5s} {ms:.3f} ms  {stream.n_bytes / ms / 1e6:8.1f} GB/s  same_as_l2={ok}", flush=True)
