"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN, "small_*.npz")))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["reads"] = str(g["reads"]).split("\n") if int(g["n_reads"]) else []
    # np.array("a\nb") of reads where trailing reads are empty strings keeps them; guard the count
    while len(g["reads"]) < int(g["n_reads"]):
        g["reads"].append("")
    g["k"] = int(g["k"])
    g["stride"] = int(g["stride"])
    g["create_all_kmers"] = bool(g["create_all_kmers"])
    g["small_ks"] = [int(s) for s in g["small_ks"]]
    g["kf_settings"] = [(float(t), None if p < 0 else float(p)) for t, p in g["kf_settings"]]
    return g


def edge_keys(ei):
    ei = np.asarray(ei)
    return ei[0].astype(np.int64) << 32 | ei[1].astype(np.int64)


def sort_edges(ei, w):
    k = edge_keys(ei)
    o = np.argsort(k, kind="stable")
    return k[o], np.asarray(w)[o]


def assert_kf_threshold_parity(ei_ours, w_ours, ei_ref, w_ref, thr, band=1e-12):
    """Threshold-bound KF parity: identical (s,t) sets outside the
    |cos - thr| <= band equality band, float32 weights bit-equal and float64
    weights within 1e-12 relative on the common edges.  Returns the band sizes."""
    ka, wa = sort_edges(ei_ours, w_ours)
    kb, wb = sort_edges(ei_ref, w_ref)
    only_a = ~np.isin(ka, kb)
    only_b = ~np.isin(kb, ka)
    assert np.all(np.abs(wa[only_a] - thr) <= band), "edge only in ours lies outside the threshold band"
    assert np.all(np.abs(wb[only_b] - thr) <= band), "edge only in reference lies outside the threshold band"
    ca, cb = wa[~only_a], wb[~only_b]
    assert np.array_equal(ka[~only_a], kb[~only_b])
    assert np.array_equal(ca.astype(np.float32).view(np.int32), cb.astype(np.float32).view(np.int32))
    if ca.size:
        assert np.max(np.abs(ca - cb) / cb) <= 1e-12
    return int(only_a.sum()), int(only_b.sum())
