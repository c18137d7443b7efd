"""Per-stage device timings of the k=8 / 1M-read build (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genomic_gnn_b200 import core
from oracle import fast

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), r

k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
reads = fast.synthetic_reads(R, 150, seed=0)
t0 = time.time(); stream = core.upload_reads(reads); torch.cuda.synchronize(); print("upload s", time.time() - t0)
lib = core._lib.load()
from genomic_gnn_b200._lib import ptr
hist = torch.empty(4 ** (k + 1), dtype=torch.int32, device="cuda"); first = torch.empty(4 ** (k + 1), dtype=torch.int64, device="cuda")
status = torch.empty(8, dtype=torch.int64, device="cuda"); br = torch.empty((1024, 4), dtype=torch.int32, device="cuda")
st = core._stream()
def f_count():
    lib.gg_count_reset(k, ptr(hist), ptr(first), ptr(status), st)
    lib.gg_count_kp1mers(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, 0, ptr(hist), ptr(br), 1024, ptr(status), st)
ms, _ = timed(f_count); print(f"count (L2 path) ms {ms:.3f}  -> {stream.n_bytes/ms/1e6:.1f} GB/s")
wsb = int(lib.gg_count_smem_workspace_bytes(stream.n_bytes, k))
if wsb:
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    def f_count_smem():
        lib.gg_count_reset(k, ptr(hist), ptr(first), ptr(status), st)
        lib.gg_count_kp1mers_smem(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, 0, ptr(hist), ptr(br), 1024, ptr(status), ptr(ws), wsb, st)
    ms, _ = timed(f_count_smem); print(f"count (smem path) ms {ms:.3f}  -> {stream.n_bytes/ms/1e6:.1f} GB/s")
def f_first():
    first.fill_(-1)
    lib.gg_first_occurrence(ptr(stream.data), stream.n_bytes, stream.fixed_len, k, 0, ptr(hist), ptr(first), ptr(status), st)
ms, _ = timed(f_first); print(f"first-occurrence ms {ms:.3f}; cursor groups {int(status[2])}")
cr = core.count_kp1mers(stream, k)
ms, db = timed(lambda: core.build_db(cr)); print(f"db build ms {ms:.3f} N={db.num_nodes} E={db.num_edges}")
for s in (2, 5):
    ms, kc = timed(lambda: core.kf_counts(db.node_codes, k, s)); print(f"counts s={s} ms {ms:.3f}")
    ms, x = timed(lambda: core.kf_features_f32(kc)); print(f"features s={s} ms {ms:.3f}")
    for impl in (["dp4a", "tcgen05"] if s == 2 else ["tcgen05x2"]):
        t0 = time.time()
        ms, kf = timed(lambda: core.kf_edges(kc.counts, kc.sqnorm, 0.8, 0.01, sq_max=kc.m ** 2, impl=impl), n=2)
        print(f"kf s={s} impl={impl} ms {ms:.3f} E={kf.edge_index.shape[1]}  pairs/s {db.num_nodes*(db.num_nodes-1)/2/ms/1e6:.1f} G")
        recs, n_recs, rowcnt, params, scored = core.kf_emit(kc.counts, kc.sqnorm, 0.8, kc.m ** 2, impl=impl, rec_capacity=1 << 25)
        ms, _ = timed(lambda: core.kf_emit(kc.counts, kc.sqnorm, 0.8, kc.m ** 2, impl=impl, rec_capacity=1 << 25), n=3)
        print(f"   emit only ms {ms:.3f}")
        ms, _ = timed(lambda: core.kf_finalize(recs, n_recs, rowcnt, params, kc.sqnorm, scored), n=3)
        print(f"   finalize only ms {ms:.3f}")
