"""Generate tests/golden/*.npz by running the REFERENCE's own functions.

Run in the dev container only (needs /root/reference):

    python -m oracle.make_golden            # small full-output fixtures
    python -m oracle.make_golden --large    # + digest-only fixtures for the BASELINE configs

Small fixtures hold complete reference outputs (node list, DB edge_index /
weight, features, KF edges).  Large fixtures hold sizes and order-independent /
order-dependent digests only, so that the GPU box (which has no reference tree)
can still check full-size runs against numbers that came from the reference.
"""
import argparse
import hashlib
import os
import sys
import time

import numpy as np

from oracle import fast, ref_import

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def edge_set_digest(ei, w32):
    """Order-independent digest of an edge set: sort by (s, t) first."""
    key = ei[0].astype(np.int64) << 32 | ei[1].astype(np.int64)
    o = np.argsort(key, kind="stable")
    return digest(key[o]), digest(np.asarray(w32, np.float32)[o])


def small_cases():
    rng = np.random.default_rng(1234)

    def rand_reads(n, lo, hi, p_n, tail):
        out = []
        for _ in range(n):
            ln = int(rng.integers(lo, hi + 1))
            s = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
            s = "".join(("N" if rng.random() < p_n else ch) for ch in s)
            if tail:
                s += "N" * int(rng.integers(0, 6))
            out.append(s)
        return out

    uniform = fast.reads_to_strings(fast.synthetic_reads(200, 150, seed=0))
    lowcx = ["A" * 40, "ACACACACACACACACAC", "AAAAAAAAAACCCCCCCCCC", "ACGT" * 9, "T" * 7, "AC", ""] * 3
    return [
        # name, reads, k, stride, create_all_kmers, small_ks, kf settings [(thr, topk)]
        ("k3_uniform", uniform, 3, 1, False, (2,), [(0.0, None), (0.8, None), (0.0, 0.01)]),
        ("k4_ragged_N", rand_reads(300, 0, 60, 0.03, True), 4, 1, False, (2, 3), [(0.0, None), (0.8, 0.01), (0.6, None)]),
        ("k5_stride2", rand_reads(300, 0, 80, 0.01, True), 5, 2, False, (2,), [(0.8, None)]),
        ("k4_allkmers", rand_reads(40, 10, 40, 0.0, False), 4, 1, True, (2,), [(0.8, None)]),
        ("k3_lowcomplexity", lowcx, 3, 1, False, (1, 2), [(0.0, None), (0.5, None)]),
    ]


def run_reference(ref, reads, k, stride, cak, small_ks, kf_settings, full=True):
    out = {}
    with ref_import.quiet():
        _, pf = ref.weighted_directed_edges(reads, k=k, stride=stride, inlcudeUNK=False)
    out["pair_u"] = np.array([p[0] for p in pf.keys()])
    out["pair_v"] = np.array([p[1] for p in pf.keys()])
    out["pair_count"] = np.array(list(pf.values()), dtype=np.int64)
    with ref_import.quiet():
        g = ref.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=cak)
    nodes = list(g.nodes())
    idx = {n: i for i, n in enumerate(nodes)}
    edges = list(g.edges(data="weight"))
    out["nodes"] = np.array(nodes)
    out["edge_index"] = np.array([[idx[u] for u, _, _ in edges], [idx[v] for _, v, _ in edges]], dtype=np.int64).reshape(2, -1)
    out["weight64"] = np.array([w for _, _, w in edges], dtype=np.float64)
    for s in small_ks:
        with ref_import.quiet():
            x = ref.node_embedding_initial(g, "kmer_frequency", k, s)
        out[f"x_s{s}"] = x
        for j, (thr, topk) in enumerate(kf_settings):
            with ref_import.quiet():
                ei, w = ref.cosine_similarity_to_edge_index_weight(np.vstack(x), threshold=thr, keep_top_k=topk)
            out[f"kf_s{s}_{j}_edge_index"] = ei
            out[f"kf_s{s}_{j}_weight"] = w
    return out


def stride_case(ref):
    """Higher-stride DB graph exactly as the builders make it (kmer_ssgnn.py:215-226): stride 2 and 3,
    edge_weight_threshold = 1/4^k, quantile keep_top_k."""
    rng = np.random.default_rng(77)
    reads = []
    for _ in range(400):
        ln = int(rng.integers(0, 70))
        s = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
        reads.append("".join(("N" if rng.random() < 0.02 else ch) for ch in s) + "N" * int(rng.integers(0, 4)))
    out = {"reads": np.array("\n".join(reads)), "n_reads": len(reads), "k": 4}
    for stride in (2, 3, 5):
        for j, (thr, topk) in enumerate([(None, None), (1.0 / 4**4, 0.3), (1.0 / 4**4, None), (None, 0.5)]):
            with ref_import.quiet():
                _, pf = ref.weighted_directed_edges(reads, k=4, stride=stride, inlcudeUNK=False)
                g = ref.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False,
                                             edge_weight_threshold=thr, keep_top_k=topk)
            nodes = list(g.nodes())
            idx = {n: i for i, n in enumerate(nodes)}
            edges = list(g.edges(data="weight"))
            tag = f"s{stride}_{j}"
            out[f"{tag}_nodes"] = np.array(nodes)
            out[f"{tag}_edge_index"] = np.array([[idx[u] for u, _, _ in edges], [idx[v] for _, v, _ in edges]],
                                                dtype=np.int64).reshape(2, -1)
            out[f"{tag}_weight64"] = np.array([w for _, _, w in edges], dtype=np.float64)
            out[f"{tag}_thr"] = -1.0 if thr is None else thr
            out[f"{tag}_topk"] = -1.0 if topk is None else topk
    return out


def large_case(ref, name, n_reads, k, small_ks, thr, topk, seed=0):
    reads_arr = fast.synthetic_reads(n_reads, 150, seed=seed)
    reads = fast.reads_to_strings(reads_arr)
    t = {}
    t0 = time.time()
    with ref_import.quiet():
        _, pf = ref.weighted_directed_edges(reads, k=k, stride=1, inlcudeUNK=False)
    t["count_s"] = time.time() - t0
    t0 = time.time()
    with ref_import.quiet():
        g = ref.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    t["db_s"] = time.time() - t0
    nodes = list(g.nodes())
    idx = {n: i for i, n in enumerate(nodes)}
    lut = np.zeros(256, np.int64)
    for i, c in enumerate(b"ACGT"):
        lut[c] = i
    codes = np.zeros(len(nodes), np.int64)
    arr = np.frombuffer("".join(nodes).encode(), np.uint8).reshape(len(nodes), k)
    for j in range(k):
        codes = codes * 4 + lut[arr[:, j]]
    edges = list(g.edges(data="weight"))
    ei = np.array([[idx[u] for u, _, _ in edges], [idx[v] for _, v, _ in edges]], dtype=np.int64)
    w32 = np.array([w for _, _, w in edges], dtype=np.float64).astype(np.float32)
    out = {
        "n_reads": n_reads, "k": k, "seed": seed, "threshold": thr,
        "keep_top_k": -1.0 if topk is None else topk,
        "num_nodes": len(nodes), "num_db_edges": ei.shape[1],
        "node_codes_sha": digest(codes), "db_edge_index_sha": digest(ei), "db_weight_sha": digest(w32),
    }
    for i, s in enumerate(small_ks):
        t0 = time.time()
        with ref_import.quiet():
            x = ref.node_embedding_initial(g, "kmer_frequency", k, s)
        t[f"feat_s{s}_s"] = time.time() - t0
        out[f"x_s{s}_sha_f32"] = digest(x.astype(np.float32))
        out[f"x_s{s}_shape"] = np.array(x.shape)
        t0 = time.time()
        with ref_import.quiet():
            kei, kw = ref.cosine_similarity_to_edge_index_weight(np.vstack(x), threshold=thr, keep_top_k=topk)
        t[f"kf_s{s}_s"] = time.time() - t0
        out[f"kf_s{s}_E"] = kei.shape[1]
        sd, wd = edge_set_digest(kei, kw.astype(np.float32))
        out[f"kf_s{s}_set_sha"] = sd
        out[f"kf_s{s}_w32_sha"] = wd
        out[f"kf_s{s}_wmin"] = float(kw.min()) if kw.shape[0] else 0.0
        out[f"kf_s{s}_n_at_wmin"] = int((kw == kw.min()).sum()) if kw.shape[0] else 0
        out[f"kf_s{s}_n_above_wmin"] = int((kw > kw.min() * (1 + 1e-9)).sum()) if kw.shape[0] else 0
        above = kw > kw.min() * (1 + 1e-9) if kw.shape[0] else np.zeros(0, bool)
        sd2, _ = edge_set_digest(kei[:, above], kw[above].astype(np.float32))
        out[f"kf_s{s}_above_set_sha"] = sd2
    for kk, vv in t.items():
        out["time_" + kk] = vv
    print(name, {kk: (vv if not isinstance(vv, np.ndarray) else vv.tolist()) for kk, vv in out.items()}, flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", action="store_true")
    ap.add_argument("--full", action="store_true", help="BASELINE configs 3 and 4 at their full 1M reads (minutes each)")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    ref = ref_import.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    if not args.only:
        for name, reads, k, stride, cak, sks, kfs in small_cases():
            out = run_reference(ref, reads, k, stride, cak, sks, kfs)
            out["reads"] = np.array("\n".join(reads))
            out["n_reads"] = len(reads)
            out["k"], out["stride"], out["create_all_kmers"] = k, stride, cak
            out["small_ks"] = np.array(sks)
            out["kf_settings"] = np.array([(t, -1.0 if p is None else p) for t, p in kfs])
            np.savez_compressed(os.path.join(GOLDEN_DIR, f"small_{name}.npz"), **out)
            print("wrote", name, {kk: getattr(vv, "shape", None) for kk, vv in out.items()})
    if not args.only or args.only == "stride":
        np.savez_compressed(os.path.join(GOLDEN_DIR, "stride_k4.npz"), **stride_case(ref))
        print("wrote stride_k4")
    if args.large:
        cases = [
            ("cfg1_k3_10k", 10_000, 3, (2,), 0.0, None),
            ("cfg2_k6_100k", 100_000, 6, (2, 5), 0.0, 0.01),
            ("cfg3_k7_100k", 100_000, 7, (2, 5), 0.8, 0.01),
            ("cfg4_k8_100k", 100_000, 8, (2, 5), 0.8, 0.01),
        ]
        if args.full:
            cases += [("cfg3_k7_1m", 1_000_000, 7, (2, 5), 0.8, 0.01), ("cfg4_k8_1m", 1_000_000, 8, (2, 5), 0.8, 0.01)]
        for name, r, k, sks, thr, topk in cases:
            if args.only and args.only != name:
                continue
            out = large_case(ref, name, r, k, sks, thr, topk)
            np.savez_compressed(os.path.join(GOLDEN_DIR, f"digest_{name}.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
