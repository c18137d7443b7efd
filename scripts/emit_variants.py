"""Where the dense-candidate epilogue of the CTA-pair Gram kernel spends its time (development aid): the same k=9,
small_k=5, threshold-0 pass with the class histogram on/off and with / without record emission."""
import sys, time
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads

k = int(sys.argv[1]) if len(sys.argv) > 1 else 9
built = core.build_graph(synthetic_reads(2_000_000, 150, seed=0), k, "5", with_kf=False, keep_features=False)
kc = built.kf_counts[0]
n = kc.counts.shape[0]
sq_max = kc.m ** 2


def timed(label, thr=0.0, **kw):
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = core.kf_emit(kc.counts, kc.sqnorm, thr, sq_max, impl="tcgen05x2", retry=False, **kw)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
    pairs = n * (n - 1) / 2
    print(f"{label:44s} {ms:8.2f} ms  {2 * 1024 * pairs / ms / 1e9:8.1f} TOP/s  found={res[1]:,} slots={res.slots:,}", flush=True)


hist = torch.zeros(core.kf_hist_bins(sq_max), dtype=torch.int64, device="cuda")
timed("thr 0.8 (sparse), no hist, records", thr=0.8, rec_capacity=1 << 22)
timed("thr 0, no hist, no records (cap 0)", rec_capacity=0)
timed("thr 0, hist (m), no records", rec_capacity=0, class_hist=hist, hist_m=kc.m)
timed("thr 0, hist (no promise), no records", rec_capacity=0, class_hist=hist, hist_m=0)
timed("thr 0, no hist, records (cap 2^31)", rec_capacity=(1 << 31) - 1 if n * n * 0.03 < (1 << 31) else 1 << 30)
timed("thr 0, hist (m), records (cap 2^27)", rec_capacity=1 << 27, class_hist=hist, hist_m=kc.m)
