"""bench.py -- k=8 graph-build throughput (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step is ONE full graph construction over one batch of synthetic reads:
counts -> first occurrences -> De Bruijn assembly -> sub-k-mer features ->
all KF edge sets.  Workload = BASELINE config 4 (k=8, small_k=2,5,
threshold 0.8, top-k 0.01, 1M synthetic 150-bp reads); with N GPUs the SAME
1M reads are split over the ranks (strong scaling).

value  reads/s with the reads already resident in HBM (CUDA events, max over ranks)
e2e    reads/s through the public API from pinned HOST reads, H2D of the reads and
       D2H of every graph tensor inside the timed region
The reference arm (--impl reference) times the oracle's faithful CPU port of the
reference (pure-Python counting, blocked float64 numpy cosine) on the host cores
of the same box, on a bounded sample scaled to the full workload.
"""
import argparse
import json
import os
import statistics
import sys
import time

if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1":
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is the reference as a user runs it
    # (numpy/BLAS on every core), so undo that before numpy loads its BLAS
    del os.environ["OMP_NUM_THREADS"]

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 256 << 20
# kernels of libgg_b200.so launched by one step (CUB sort/scan kernels and memsets not counted):
# K1 partition+bucket count (k = 7, 8; one counting kernel otherwise) + pending + first = 4, K2 = 6,
# per small_k: K3 counts+features = 2 and either emit + place + expand = 3 or, for rows of <= 16 counts on >= 32768 nodes
# (duplicate-row path), row keys + heads + class starts + class scoring + count sweep + place sweep + expand = 7
def our_launches_per_step(small_ks, k, n_nodes):
    total = (4 if k in (7, 8) else 3) + 6
    for s in small_ks:
        total += 2 + (7 if (4**s <= 16 and n_nodes >= 32768) else 3)
    return total


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--small-k", default="2,5")
    ap.add_argument("--threshold", type=float, default=0.8)
    ap.add_argument("--top-k", type=float, default=0.01)
    ap.add_argument("--kf-impl", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end loop")
    ap.add_argument("--no-probes", action="store_true", help="skip the hardware-ceiling probes and the launch census")
    ap.add_argument("--kf-output", default="auto", choices=["auto", "gathered", "sharded", "discard"],
                    help="gathered: every rank ends with every KF edge (configs 1-4); sharded: every rank keeps the edges "
                         "of its own row blocks (config 5: 1.1e10 edges per set fit no single device); discard: sharded, "
                         "and a binding top-n recycles one chunk buffer (the result exceeds this many GPUs' memory); "
                         "auto: gathered up to k=8, else sharded when the shards fit, else discard")
    ap.add_argument("--chunk-records", type=int, default=0, help="packed words per top-n emission chunk (0: 2^27)")
    ap.add_argument("--nccl-only", action="store_true", help="keep every exchange on NCCL (no NVLink peer-memory pushes)")
    ap.add_argument("--debug-no-flush", action="store_true")
    ap.add_argument("--debug-no-clocks", action="store_true")
    ap.add_argument("--debug-timer-in-headline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    name = {10: "config5", 8: "config4", 7: "config3", 6: "config2", 3: "config1"}.get(a.k, "custom")  # BASELINE.json configs by k
    return (f"{name}: k={a.k}, small_k={a.small_k}, KF threshold={a.threshold}, top_k={a.top_k}, "
            f"{a.reads} synthetic {a.read_len}-bp reads (uniform ACGT, seed 0)")


def metric_name(a):
    return f"k={a.k} graph-build reads/sec"


def bench_config(a):
    """The `config` object both arms print, key for key."""
    return {"workload": workload_name(a)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p["bf16_tflops"]), "source": "measured"}
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback"}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sample(a, reads_arr, count_reads=20_000, kf_row_blocks=1, feat_nodes=8192, kf_nodes_cap=65_536):
    """Bounded sample of the reference's CPU path, scaled to the full workload.
    Returns (estimated seconds for the full workload, description, detail dict)."""
    from oracle import fast, faithful

    k, n_total = a.k, a.reads
    count_reads = min(count_reads, n_total)
    sample = fast.reads_to_strings(reads_arr[:count_reads])
    t0 = time.perf_counter()
    _, pf = faithful.weighted_directed_edges(sample, k=k, stride=1, inlcudeUNK=False)
    t_count = (time.perf_counter() - t0) * (n_total / count_reads)
    t0 = time.perf_counter()
    g = faithful.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False)
    t_db = time.perf_counter() - t0
    nodes = g.nodes()
    n_nodes = len(nodes)
    codes = None
    detail = {"count_s_scaled": t_count, "db_s": t_db, "n_nodes": n_nodes}
    t_feat = t_kf = 0.0
    # the KF sample scores `rb` row blocks against the first `n_kf` nodes (all of them up to 65,536: beyond that the
    # float64 feature matrix alone is gigabytes) and is scaled by block pairs to the full node count
    n_full = max(n_nodes, min(4**k, (n_total * max(a.read_len - k, 1))))   # every k-mer occurs at BASELINE read counts
    n_kf = min(n_nodes, kf_nodes_cap)
    nb_full, nb = (n_full + 1023) // 1024, (n_kf + 1023) // 1024
    rb = min(kf_row_blocks, nb)
    pair_scale = (nb_full * (nb_full + 1) / 2) / sum(nb - i for i in range(rb))
    for s in (int(x) for x in a.small_k.split(",")):
        sub = nodes[: min(feat_nodes, n_nodes)]
        t0 = time.perf_counter()
        faithful.node_embedding_initial(sub, "kmer_frequency", k, s)
        tf = (time.perf_counter() - t0) * (n_full / len(sub))
        if codes is None:
            lut = np.zeros(256, np.int64)
            lut[np.frombuffer(b"ACGT", np.uint8)] = np.arange(4)
            arr = np.frombuffer("".join(nodes).encode(), np.uint8).reshape(n_nodes, k)
            codes = np.zeros(n_nodes, np.int64)
            for j in range(k):
                codes = codes * 4 + lut[arr[:, j]]
        x, _ = fast.features_from_counts(fast.subkmer_counts(codes[:n_kf], k, s), k, s)  # untimed: input of the timed KF sample
        t0 = time.perf_counter()
        faithful.cosine_similarity_to_edge_index_weight(x, threshold=a.threshold, keep_top_k=None, row_blocks=rb)
        tk = (time.perf_counter() - t0) * pair_scale
        detail[f"feat_s{s}_s_scaled"] = tf
        detail[f"kf_s{s}_s_scaled"] = tk
        t_feat += tf
        t_kf += tk
    total = t_count + t_db + t_feat + t_kf
    desc = (f"counting: first {count_reads} reads x{n_total / count_reads:.0f}; DB assembly of the sample; features: first "
            f"{min(feat_nodes, n_nodes)} nodes x{n_full / min(feat_nodes, n_nodes):.0f}; KF: {rb} row block(s) against "
            f"{n_kf} nodes x{pair_scale:.1f} by block pairs to {n_full} nodes; counting/features are single-thread Python, "
            f"KF is numpy/BLAS")
    return total, desc, detail


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


def reference_full_pass(a, ref, reads_arr):
    """ONE un-extrapolated pass of the reference's own functions (imported unmodified, oracle/ref_import.py) over the
    whole workload, exactly the calls of kmer_ssgnn_miniBatch.py:174-199 + dataloader.py:106-137."""
    from oracle import fast
    from oracle.ref_import import quiet

    detail = {}
    seqs = fast.reads_to_strings(reads_arr)
    with quiet():
        t0 = time.perf_counter()
        _, pf = ref.weighted_directed_edges(seqs, k=a.k, stride=1, inlcudeUNK=False, disable_tqdm=True)
        detail["count_s"] = time.perf_counter() - t0
        t1 = time.perf_counter()
        g = ref.build_deBruijn_graph(pf, normalise=True, remove_N=True, create_all_kmers=False, disable_tqdm=True)
        detail["db_s"] = time.perf_counter() - t1
        labels = []
        for s_ in (int(x) for x in a.small_k.split(",")):
            t1 = time.perf_counter()
            labels.append(ref.node_embedding_initial(g, method="kmer_frequency", k=a.k, small_k=s_))
            detail[f"feat_s{s_}_s"] = time.perf_counter() - t1
        n_edges = []
        for s_, lab in zip((int(x) for x in a.small_k.split(",")), labels):
            t1 = time.perf_counter()
            ei, _ = ref.cosine_similarity_to_edge_index_weight(np.vstack(lab), threshold=a.threshold, keep_top_k=a.top_k)
            detail[f"kf_s{s_}_s"] = time.perf_counter() - t1
            n_edges.append(int(ei.shape[1]))
        total = time.perf_counter() - t0
    detail.update({"n_nodes": g.number_of_nodes(), "kf_edges": n_edges})
    return total, detail


def run_reference_arm(a, rank):
    """--impl reference: the reference's CPU graph construction on this box's host cores, rank 0 only.
    With the reference's modules staged (oracle/_ref, or /root/reference in the dev container) and k <= 8 this is ONE
    full pass of the reference itself -- reported as steps=1 / warmup=0 whatever was requested, because a pass takes
    minutes; otherwise (k >= 9: the reference needs hours and 8 GB for X) a bounded, scaled sample of the oracle port."""
    if rank != 0:
        return
    from genomic_gnn_b200.synth import synthetic_reads
    from oracle import ref_import

    try:  # torchrun exports OMP_NUM_THREADS=1; the reference's numpy Gram uses every core when run as the user would
        from threadpoolctl import threadpool_limits

        limiter = threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        limiter = None
    requested = {"steps": a.steps, "warmup": a.warmup}
    if ref_import.available() and a.k <= 8:
        ref = ref_import.load()
        reads = synthetic_reads(a.reads, a.read_len, seed=0)
        sec, detail = reference_full_pass(a, ref, reads)
        kind, steps, warmup = "reference", 1, 0
        desc = (f"the reference's own functions ({os.path.relpath(ref.root, ROOT) if ref.root.startswith(ROOT) else ref.root}), "
                f"one full un-extrapolated pass over all {a.reads} reads; counting / assembly / features are "
                f"single-thread Python, the KF Gram is numpy/BLAS on {blas_threads()} threads")
    else:
        reads = synthetic_reads(min(a.reads, 50_000), a.read_len, seed=0)
        vals, desc, detail = [], "", {}
        for i in range(a.warmup + a.steps):
            total_s, desc, detail = cpu_reference_sample(a, reads)
            if i >= a.warmup:
                vals.append(total_s)
        sec = statistics.mean(vals)
        kind, steps, warmup = "port", a.steps, a.warmup
        detail["extrapolated"] = True
    value = a.reads / sec
    line = {
        "impl": "reference", "metric": metric_name(a), "value": value, "unit": "reads/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "requested": requested, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(a),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": blas_threads(), "kind": kind, "sample": desc,
                         "host_cpus": os.cpu_count(), "detail": detail},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    del limiter
    print(json.dumps(line), flush=True)


def census_launches(step, profile=True):
    """Kernel launches of ONE untimed step, counted by torch.profiler (CUPTI): kernels of libgg_b200.so ("own"),
    CUB sort / scan kernels instantiated inside it, torch's own elementwise / copy kernels, NCCL.  The step runs exactly
    once whatever happens to the profiler: other ranks are running it too."""
    import torch

    ran = [False]

    def once():
        if not ran[0]:
            ran[0] = True
            step()
            torch.cuda.synchronize()

    if not profile:      # a rank that only has to take part in the step's exchanges
        once()
        return None
    try:
        from torch.profiler import ProfilerActivity, profile as torch_profile

        torch.cuda.synchronize()
        with torch_profile(activities=[ProfilerActivity.CUDA]) as prof:
            once()
        out = {"own": 0, "cub": 0, "torch": 0, "nccl": 0, "memcpy_memset": 0}
        names = {}
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            name = ev.name
            low = name.lower()
            if low.startswith("memcpy") or low.startswith("memset"):
                out["memcpy_memset"] += 1
            elif "nccl" in low:
                out["nccl"] += 1
            elif "cub::" in name:
                out["cub"] += 1
            elif "at::" in name or "at_cuda" in name:
                out["torch"] += 1
            else:
                out["own"] += 1
                short = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0].split("::")[-1]
                names[short] = names.get(short, 0) + 1
        out["own_kernels"] = names
        out["source"] = "torch.profiler (CUPTI), one untimed step"
        return out
    except Exception as exc:  # noqa: BLE001
        sys.stderr.write(f"launch census unavailable: {exc}\n")
        once()
        return None


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions.  NVML is read in-process from a thread
    (a resident `nvidia-smi -lms` loop was measured to stall kernel launches for milliseconds per poll)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device_index, period_s=0.02):
        import threading

        import torch

        self.sm, self.reasons, self.max_sm, self.err = [], set(), None, None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception as exc:  # noqa: BLE001
            self.err = f"NVML unavailable: {exc}"
            self.thread = None
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception as exc:  # noqa: BLE001
                    self.err = str(exc)
                self._stop.wait(period_s)

        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err]}
        self._stop.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "NVML, 20 ms period"}


# --------------------------------------------------------------------------- our arm
def builder_e2e(a, reads_host, steps):
    """The call the reference makes: KmerSsgnnMb.create_graphs(self, k) (kmer_ssgnn_miniBatch.py:160-285) with
    ``self.ssgnn_dataset`` a Python list of read strings (kmer_ssgnn.py:56-57), timed host to host -- list[str] in,
    HeteroData-shaped CPU tensors + loader out -- with the share spent marshalling the strings."""
    import types

    import torch

    import genomic_gnn_b200 as gg

    arr = reads_host.numpy()
    seqs = [bytes(r).decode("ascii") for r in arr]
    me = types.SimpleNamespace(ssgnn_dataset=seqs, strides=[1], small_k=a.small_k, create_all_kmers=False,
                               disable_tqdm=True, representation_ss_initial_labels="kmer_frequency", batch_size=256,
                               kmer_frequency_loss_weight=1.0, edit_distance_loss_weight=0.0,
                               edges_threshold=a.threshold, edges_keep_top_k=a.top_k, layers_config=["32_KF0"],
                               representation_ss_last_layer_edge_type="KF0", ss_task="CL", faiss_ann=False)
    times, marshal = [], []
    for i in range(3 + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gg.create_graphs_mini_batch(me, a.k)
        torch.cuda.synchronize()
        if i >= 3:
            times.append((time.perf_counter() - t0) * 1e3)
            marshal.append(me.build_info["marshal_ms"])
    ms = statistics.mean(times)
    return {"value": a.reads / (ms * 1e-3), "unit": "reads/s", "ms_per_step": ms, "marshal_ms": statistics.mean(marshal),
            "marshal_share": statistics.mean(marshal) / ms, "input": "list[str]", "timer": "host wall clock, CUDA synchronised",
            "call": "genomic_gnn_b200.create_graphs_mini_batch(self, k)", "steps": steps,
            "kf_edges": [int(me.torch_graph["node", f"KF{i}", "node"].edge_index.shape[1])
                         for i in range(len(a.small_k.split(",")))]}


def kf_output_mode(a, world):
    """Where the KF edge sets end up (see --kf-output)."""
    if a.kf_output != "auto":
        return a.kf_output
    if a.k <= 8 or a.top_k is None:
        return "gathered"
    n = 4**a.k
    per_rank = len(a.small_k.split(",")) * int(n * n * a.top_k) * 20 / world + 24e9   # kept edges + chunk staging
    return "sharded" if per_rank < 0.8 * 180e9 else "discard"


def run_ours(a):
    import torch

    from genomic_gnn_b200 import core, dist as ggdist, hostshare
    from genomic_gnn_b200.synth import synthetic_reads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist

        import datetime

        # a rank that dies must take the others down in minutes, not in NCCL's default half hour
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
        comm = ggdist.TorchComm()
        if not a.nccl_only:   # count pieces, candidate counts and packed KF edges over NVLink peer memory
            comm.enable_peer_exchange((1 << 30) if a.k <= 8 else (256 << 20))
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world}"

    reads_all = synthetic_reads(a.reads, a.read_len, seed=0)
    lo, hi = ggdist.shard_bounds(a.reads, world, rank)
    reads_host = torch.from_numpy(reads_all[lo:hi].copy()).pin_memory()
    del reads_all
    pos_base = lo * a.read_len
    resident = core.upload_reads(reads_host, pos_base=pos_base)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    small_ks = [int(x) for x in a.small_k.split(",")]
    n_small = len(small_ks)
    mode = kf_output_mode(a, world)
    big = a.k >= 9                       # config 5: no float32 feature matrices (4 GB per wide set), sharded KF output
    kw = dict(k=a.k, small_k=a.small_k, edges_threshold=a.threshold, edges_keep_top_k=a.top_k, kf_impl=a.kf_impl,
              comm=comm, kf_gather=mode == "gathered", kf_sink="discard" if mode == "discard" else "device",
              kf_chunk_records=a.chunk_records or None, keep_features=not big)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return core.build_graph(resident, **kw)

    # end to end: pinned host reads -> H2D -> build -> graph tensors in host memory.  One GPU: the public host-to-host
    # call (copies overlapped with the kernels).  Several GPUs: every rank uploads its reads, builds with the KF sets
    # left sharded and copies ITS shard into a host arena shared by all ranks -- no edge all-gather, all PCIe links busy.
    # Config 5: the KF edge sets stay sharded on the devices (BASELINE: "output left sharded"; 220 GB per set do not
    # belong in host memory); what comes down is the De Bruijn graph.
    arena = None
    if not a.no_e2e and world > 1 and not big:
        n_nodes_max = min(4**a.k, a.reads * a.read_len)
        budgets = [int(n_nodes_max**2 * a.top_k) if a.top_k else n_nodes_max * 64 for _ in small_ks]
        arena = hostshare.SharedHostArena(comm, hostshare.arena_bytes_for(n_nodes_max, a.k, small_ks, budgets))
    kw_e2e = dict(kw, kf_gather=False) if world > 1 else kw

    def step_e2e():
        if world == 1 and not big:
            g = core.build_graph_to_host(reads_host, a.k, a.small_k, edges_threshold=a.threshold,
                                         edges_keep_top_k=a.top_k, kf_impl=a.kf_impl)
            return g, g.d2h_bytes
        stream = core.upload_reads(reads_host, pos_base=pos_base)     # H2D of this rank's reads (pinned source)
        built = core.build_graph(stream, **kw_e2e)
        if big:
            nbytes = 0
            if rank == 0:
                nbytes = core.graph_to_host(core.BuiltGraph(built.k, built.db, [], [], [])).d2h_bytes
            summary = [kf.per_rank for kf in built.kf]                # the step's result for the KF sets: edges per rank
            torch.cuda.current_stream().synchronize()
            return (built, summary), nbytes
        host = hostshare.deliver_graph(built, arena, comm)
        return host, host.d2h_bytes

    def timed_loop(fn, steps, warmup, timer=None):
        last = None
        for _ in range(warmup):
            last = None          # release the previous result first: config 5 keeps tens of GB per step
            last = fn()
        barrier()
        core.set_timer(timer)
        total_ms = 0.0
        for _ in range(steps):
            last = None
            if not a.debug_no_flush:
                flush.fill_(1)                                        # evict L2 between timed iterations
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            last = fn()
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        core.set_timer(None)
        barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), last

    # measured ceilings for the stages that are not HBM-bound, in this process and at these clocks (rank 0)
    probes = {}
    if rank == 0 and not a.no_probes:
        from genomic_gnn_b200 import probes as ggprobes

        probes = {"tensor_i8": ggprobes.tensor_i8_rate(), "smem_atomics": ggprobes.smem_atomic_rate()}
    clocks = ClockSampler(local_rank) if rank == 0 and not a.debug_no_clocks else None
    timer = core.StageTimer()
    # headline: K clean steps; the per-stage CUDA events (roofline numbers) come from a separate short pass so that
    # their host-side cost cannot leak into the headline
    total_ms, built = timed_loop(step_resident, a.steps, a.warmup, timer if a.debug_timer_in_headline else None)
    stage_steps = a.steps
    if not a.debug_timer_in_headline:
        kf_meta = [(kf.n_edges, kf.per_rank, kf.n_candidates, kf.cutoff, kf.tie_kept, kf.tie_total, kf.resident) for kf in built.kf]
        if big:
            built.kf = []        # free the kept shards before the next build allocates its own
        stage_steps = min(a.steps, 2 if big else 5)
        _, built2 = timed_loop(step_resident, stage_steps, 0 if big else 1, timer)
        if not big:
            built = built2
        del built2
    else:
        kf_meta = [(kf.n_edges, kf.per_rank, kf.n_candidates, kf.cutoff, kf.tie_kept, kf.tie_total, kf.resident) for kf in built.kf]
    # a stage may launch several times per step (top-n: scoring pass + one emission per chunk): sum, then per step
    stage_ms = {k_: sum(v) / stage_steps for k_, v in timer.summary().items()}
    stage_launches = {k_: len(v) / stage_steps for k_, v in timer.summary().items()}
    n_nodes, db_edges = built.db.num_nodes, built.db.num_edges
    # every rank runs the step (it holds collectives); rank 0 profiles it
    launches = census_launches(step_resident, rank == 0) if not a.no_probes and not big else None
    if big:
        built = None
    if a.no_e2e:
        e2e_ms, d2h_bytes = float("nan"), 0
    else:
        e2e_ms, (_, d2h_bytes) = timed_loop(step_e2e, a.steps if not big else min(a.steps, 2), max(a.warmup, 3) if not big else 1)
        e2e_ms = e2e_ms / (a.steps if not big else min(a.steps, 2))
    if world > 1:                        # bytes delivered by ALL ranks
        import torch.distributed as dist

        t = torch.tensor([float(d2h_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        d2h_bytes = int(t.item())
    clock_info = clocks.stop() if clocks else None
    if arena is not None:
        arena.close()

    ms_per_step = total_ms / a.steps
    value = a.reads / (ms_per_step * 1e-3)
    kf_ms = sum(v for k_, v in stage_ms.items() if k_.startswith(("k4_", "k5_", "k6_", "c2_", "c5_", "re_")))
    kf_edges = sum(m[0] for m in kf_meta)

    peaks = measured_peaks()
    local_bytes = (hi - lo) * a.read_len + 4 ** (a.k + 1) * 4
    rooflines = {}
    if "k1_count" in stage_ms:
        ach = local_bytes / (stage_ms["k1_count"] * 1e-3) / 1e9
        rooflines["k1_count"] = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm_gbs"], "traffic": None, "ms": stage_ms["k1_count"],
                                 "algorithmic_bytes": local_bytes, "peak_source": peaks["source"]}
        windows = (hi - lo) * max(a.read_len - a.k, 0)
        rate = windows / (stage_ms["k1_count"] * 1e-3)
        if "smem_atomics" in probes and a.k <= 8:
            # secondary ceiling: one histogram update per (k+1)-mer window, against the measured shared-memory atomic
            # rate -- the bound of a 4^(k+1)-bin histogram over near-uniform keys on this GPU
            rooflines["k1_count"]["atomic_ceiling"] = {
                "updates": windows, "achieved_updates_per_s": rate, "peak_updates_per_s": probes["smem_atomics"]["updates_per_s"],
                "frac": rate / probes["smem_atomics"]["updates_per_s"],
                "how": "gg_probe_smem_atomics: uniform 32-bit ATOMS into a 32,768-bin table per CTA, 1024 threads x all SMs"}
        elif a.k >= 9:
            rooflines["k1_count"]["atomic_ceiling"] = {
                "updates": windows, "achieved_updates_per_s": rate,
                "note": "4^(k+1) >= 2^20 bins: one RED per window into the L2-resident table (the partition path stops at k=8)"}
    for i, s in enumerate(small_ks):
        d = 4**s
        pairs = n_nodes * (n_nodes - 1) / 2 / world  # strict upper triangle, this rank's share (balanced shards)
        flops = 2 * d * pairs
        key = next((c for c in (f"k4_emit_d{d}", f"k4_emit_d{max(d, 32)}") if c in stage_ms), None)  # narrow rows are padded
        if key is not None:
            # a binding top-n scores the shard twice on the tensor path: once for the class table, once chunk by chunk
            # against the cut-off -- both passes are algorithmic work of this method, both are credited
            re_ms = stage_ms.get("re_" + key, 0.0)
            n_emit = stage_launches[key] + stage_launches.get("re_" + key, 0)
            passes = 2 if re_ms > 0 else 1
            ms_emit = stage_ms[key] + re_ms
            ach = passes * flops / (ms_emit * 1e-3) / 1e12
            rooflines[key] = {"bound": "tensor" if d >= 128 else "alu", "achieved": ach, "peak": peaks["tflops"],
                              "unit": "TFLOP/s", "frac": ach / peaks["tflops"], "traffic": None, "ms": ms_emit,
                              "algorithmic_flops": passes * flops, "gram_passes": passes, "launches_per_step": n_emit,
                              "first_pass_ms": stage_ms[key], "re_emission_ms": re_ms,
                              "peak_source": peaks["source"] + " bf16 burst",
                              "pairs_per_s": passes * pairs / (ms_emit * 1e-3)}
            if d >= 128 and "tensor_i8" in probes:
                # the kernel computes in kind::i8 (2x the bf16 rate): quote it against the int8 MMA rate measured here,
                # in this process, and keep the bf16 figure of MEASURED_PEAKS.json as a note
                i8 = probes["tensor_i8"]["tops"]
                rooflines[key].update({"peak": i8, "frac": ach / i8, "peak_source": "measured in-process: " + probes["tensor_i8"]["how"],
                                       "peak_bf16_note": {"peak": peaks["tflops"], "frac_of_bf16": ach / peaks["tflops"],
                                                          "source": peaks["source"] + " bf16 burst (MEASURED_PEAKS.json)"}})
        elif f"k4_score_d{d}" in stage_ms:
            # duplicate-row path: classes + class-pair scoring + per-class count and placement stand in for the emit
            # and place kernels.  ALU / issue-bound: a tensor fraction means nothing here.  Quote the edges it writes
            # against the HBM write floor instead: 8 bytes per packed edge (place) + 20 per expanded edge.
            parts = [f"k4_classes_d{d}", f"k4_score_d{d}", f"k4_count_d{d}", f"k6_place_d{d}"]
            ms = sum(stage_ms.get(p_, 0.0) for p_ in parts)
            e_local = (kf_meta[i][1][rank] if (kf_meta[i][1] and mode != "gathered") else kf_meta[i][0] / world)
            ms_all = ms + stage_ms.get(f"k6_expand_d{d}", 0.0) + (stage_ms.get("k5_select", 0.0) if kf_meta[i][3] is not None else 0.0)
            out_bytes = e_local * 28.0
            ach_gbs = out_bytes / (ms_all * 1e-3) / 1e9
            rooflines[f"k4_dedup_d{d}"] = {"bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                           "frac": ach_gbs / peaks["hbm_gbs"], "traffic": None, "ms": ms_all,
                                           "algorithmic_bytes": out_bytes, "peak_source": peaks["source"],
                                           "edges_per_s": e_local / (ms_all * 1e-3), "pairs_per_s": pairs / (ms * 1e-3),
                                           "stages": parts + [f"k6_expand_d{d}"],
                                           "note": "issue-bound sparse class sweeps; floor = edge bytes written once"}
    dominant = max(rooflines, key=lambda k_: rooflines[k_]["ms"]) if rooflines else None
    # the per-row kNN mode of the KF edges (--representation_ss_faiss_ann: kf_faiss_to_edge_index_weight, k = int(p N)
    # neighbours per node) is not part of the step; it is quoted beside it for the widest set (metric "IP"): the default
    # path for rows this sparse (an inverted-index join, shared-memory bound) and the two Gram passes on the tensor kernel
    knn_line = None
    if rank == 0 and world == 1 and not big and not a.no_probes and a.top_k:
        kc = max(built.kf_counts, key=lambda kc_: int(kc_.counts.shape[1]))
        d_knn, k_nn = int(kc.counts.shape[1]), int(a.top_k * n_nodes)
        if d_knn >= 128 and 1 <= k_nn < n_nodes:
            def time_knn(impl, reps=3):
                for _ in range(2):
                    core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, "IP", impl=impl)
                ktimer = core.StageTimer()
                core.set_timer(ktimer)
                for _ in range(reps):
                    flush.fill_(1)
                    res = core.kf_knn_edges(kc.counts, kc.sqnorm, k_nn, "IP", impl=impl)
                core.set_timer(None)
                kms = {k_: sum(v) / reps for k_, v in ktimer.summary().items()}
                return kms, sum(kms.values()), int(res.edge_index.shape[1])

            kms, total_ms, n_e = time_knn("auto")
            knn_line = {"impl": "auto: inverted-index join (knn_sparse_kernel<1|2>) + plan + reorder", "ms": total_ms,
                        "stage_ms": kms, "k": k_nn, "metric": "IP", "nodes": n_nodes, "d": d_knn, "edges": n_e,
                        "pairs_per_s": n_nodes * (n_nodes - 1) / (total_ms * 1e-3), "edges_per_s": n_e / (total_ms * 1e-3),
                        "bound": "shared memory / issue (sparse rows: only pairs with a non-zero dot are formed)"}
            try:
                tms, t_total, _ = time_knn("tcgen05")
                gram_ms = tms.get(f"knn_count_d{d_knn}", 0.0) + tms.get(f"knn_write_d{d_knn}", 0.0)
                kflops = 2 * 2.0 * d_knn * n_nodes * n_nodes
                ach = kflops / (gram_ms * 1e-3) / 1e12
                tline = {"ms": t_total, "stage_ms": tms, "bound": "tensor", "achieved": ach, "unit": "TFLOP/s",
                         "algorithmic_flops": kflops, "gram_passes": 2,
                         "kernel": "knn_tc_pair_kernel<1|2> (tcgen05.mma.cta_group::2.kind::i8) + plan + reorder",
                         "note": "epilogue-bound: per-pair work on the non-zero dots of every row (DESIGN.md section 7b)"}
                if "tensor_i8" in probes:
                    tline.update({"peak": probes["tensor_i8"]["tops"], "frac": ach / probes["tensor_i8"]["tops"],
                                  "peak_source": "measured in-process: " + probes["tensor_i8"]["how"]})
                knn_line["tcgen05"] = tline
            except Exception as exc:  # noqa: BLE001
                knn_line["tcgen05"] = {"unavailable": str(exc)}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if rooflines and os.path.exists(traffic_file) and not big:  # per-launch DRAM bytes from the committed ncu --set full capture
        with open(traffic_file) as fh:
            measured = json.load(fh)
        for key_, r_ in rooflines.items():
            r_["traffic"] = measured.get(key_)

    static_launches = our_launches_per_step(small_ks, a.k, n_nodes)
    line = {
        "metric": metric_name(a), "value": value, "unit": "reads/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u8 counts, int32/fp32 accumulate, f64 weights",
        "data": "synthetic",
        "config": bench_config(a),
        "run": {"parallelism": f"reads and KF row blocks sharded over {world} GPU(s)", "kf_output": mode,
                "exchange": ("none" if world == 1 else
                             ("gg_peer_push over NVLink peer memory (symmetric memory), NCCL for the rest"
                              if comm.peers is not None else "NCCL (" + getattr(comm, "peer_error", "--nccl-only") + ")")),
                "l2": "256 MiB buffer written between timed iterations", "kf_impl": a.kf_impl},
        "kf_edges_per_s": kf_edges / (kf_ms * 1e-3) if kf_ms else None,
        "kf_pairs_per_s": n_small * n_nodes * (n_nodes - 1) / 2 / (kf_ms * 1e-3) if kf_ms else None,
        "kf_edges": kf_edges, "nodes": n_nodes, "db_edges": db_edges,
        "kf_sets": [{"small_k": s_, "edges": m[0], "edges_per_rank": m[1], "candidates": m[2], "cutoff": m[3],
                     "cutoff_class_kept": m[4], "cutoff_class_size": m[5], "resident": m[6]}
                    for s_, m in zip(small_ks, kf_meta)],
        "stage_ms": stage_ms, "stage_launches_per_step": stage_launches,
        "e2e": {"value": a.reads / (e2e_ms * 1e-3), "unit": "reads/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": a.reads * a.read_len, "d2h_bytes_per_step": d2h_bytes,
                "delivery": ("KF edge sets stay sharded on the devices (config 5); De Bruijn graph to pinned host memory" if big
                             else ("every rank copies its shard into one page-locked shared-memory arena" if world > 1
                                   else "build_graph_to_host: copies overlapped with the kernels"))},
        "gpu_launches": (launches["own"] if launches else static_launches) * a.steps,
        "gpu_launches_per_step": launches or {"own": static_launches, "source": "static count (no profiler pass)"},
        "probes": probes,
        "roofline": {"kernel": dominant, **rooflines[dominant]} if dominant else None,
        "rooflines": rooflines,
        "knn_mode": knn_line,
        "clocks": clock_info,
    }
    if rank == 0 and world == 1 and not big and not a.no_e2e:
        line["e2e_builder"] = builder_e2e(a, reads_host, min(a.steps, 5))
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sec, desc, detail = cpu_reference_sample(a, synthetic_reads(min(a.reads, 50_000), a.read_len, seed=0))
        line["cpu_baseline"] = {"value": a.reads / sec, "unit": "reads/s", "cores": blas_threads(), "kind": "port",
                                "sample": desc, "host_cpus": os.cpu_count(), "est_seconds_full_workload": sec,
                                "detail": detail}
    if world > 1:   # what every rank measured: the slowest rank of a stage is what the step waits for
        import torch.distributed as dist

        per_rank = [None] * world
        dist.all_gather_object(per_rank, {k_: round(v, 4) for k_, v in stage_ms.items()})
        line["stage_ms_per_rank"] = per_rank
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a, int(os.environ.get("RANK", "0")))
        return
    run_ours(a)


if __name__ == "__main__":
    main()
