"""k=10 (1,048,576 nodes) through the device pipeline: cross-checks at the largest node count the KF kernels
support (development aid; the pytest version is tests/test_gpu_parity.py::test_k10_million_nodes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

thr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.95
t0 = time.time()
reads = synthetic_reads(2_000_000, 150, seed=0)
print("reads", time.time() - t0, flush=True)
torch.cuda.synchronize(); t0 = time.time()
built = core.build_graph(reads, 10, "2,5", with_kf=False)
torch.cuda.synchronize()
print("count+db+features s", time.time() - t0, "N", built.db.num_nodes, "E_DB", built.db.num_edges, flush=True)
print("mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
kc2, kc5 = built.kf_counts
for impl in ("dedup", "tcgen05"):
    torch.cuda.synchronize(); t0 = time.time()
    kf = core.kf_edges(kc2.counts, kc2.sqnorm, thr, None, sq_max=kc2.m ** 2, impl=impl)
    torch.cuda.synchronize()
    print(impl, "s=2 edges", kf.edge_index.shape[1], "s", round(time.time() - t0, 3), "mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
    if impl == "dedup":
        ref = kf
    else:
        assert torch.equal(ref.edge_index, kf.edge_index) and torch.equal(ref.weight64, kf.weight64)
        print("dedup == tcgen05 at N = 2^20", flush=True)
del ref, kf
torch.cuda.synchronize(); t0 = time.time()
kf5 = core.kf_edges(kc5.counts, kc5.sqnorm, 0.8, None, sq_max=kc5.m ** 2)
torch.cuda.synchronize()
dt = time.time() - t0
n = built.db.num_nodes
print("s=5 edges", kf5.edge_index.shape[1], "s", round(dt, 3), "PFLOP/s", 1024 * n * (n - 1) / dt / 1e15, flush=True)
# sample rows against numpy
rng = np.random.default_rng(0)
rows = np.sort(rng.choice(n, 48, replace=False))
c = kc5.counts.cpu().numpy()
sq = kc5.sqnorm.cpu().numpy().astype(np.int64)
dots = c[rows].astype(np.int32) @ c.T.astype(np.int32)          # [48, n]
cos = dots / np.sqrt(sq[rows][:, None] * sq[None, :])
ei = kf5.edge_index.cpu().numpy()
for r, row in zip(rows, cos):
    want = np.nonzero((row > 0.8) & (np.arange(n) > r))[0]
    got = np.sort(ei[1][ei[0] == r])
    assert np.array_equal(want, got), (r, want[:5], got[:5])
print("48 sampled rows of the s=5 set match numpy", flush=True)
del kf5, ei, cos, dots
# steady-state timings (second calls)
for name, fn in (("k1+k2+k3", lambda: core.build_graph(reads_dev, 10, "2,5", with_kf=False)),
                 ("kf s=2 dedup", lambda: core.kf_edges(kc2.counts, kc2.sqnorm, thr, None, sq_max=kc2.m ** 2)),
                 ("kf s=5 tensor", lambda: core.kf_edges(kc5.counts, kc5.sqnorm, 0.8, None, sq_max=kc5.m ** 2))):
    if name == "k1+k2+k3":
        reads_dev = core.upload_reads(reads)
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timer = core.StageTimer(); core.set_timer(timer)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    core.set_timer(None)
    print(name, "ms", round(a.elapsed_time(b), 3), {k_: round(v[0], 3) for k_, v in timer.summary().items()}, flush=True)
