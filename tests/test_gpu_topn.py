"""Global top-n of the KF edges WITHOUT materialising the candidates (gnn_utils.py:311-317 at sizes where
`np.argpartition` over every surviving pair is out of reach: BASELINE config 5 has 5.5e11 of them).

The score classes are counted while scoring (CTA-pair tensor kernel: in its epilogue; duplicate-row path: class pairs
weighted by member counts), the cut-off comes from that table, and the shard is re-scored in row-block chunks against
the cut-off.  Checked here, at sizes the oracle finishes in seconds: identical edges, order and weights to
oracle/fast.py's statement of the same policy, for every chunk size, with and without the row-sum promise, and for
sharded runs on world sizes 2 and 3 (threads standing in for ranks on the one GPU of the test box; the NCCL run is
tests/test_multi_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import fast

from ._threadcomm import ThreadWorld

pytestmark = pytest.mark.gpu


def _counts(k, s, n_reads, seed):
    from genomic_gnn_b200 import core

    reads = fast.synthetic_reads(n_reads, 150, seed=seed)
    stream = core.upload_reads(reads)
    db = core.build_db(core.count_kp1mers(stream, k))
    kc = core.kf_counts(db.node_codes, k, s)
    return kc, kc.counts.cpu().numpy()


def _check(kf, want_ei, want_w):
    assert np.array_equal(kf.edge_index.cpu().numpy(), want_ei)
    assert np.array_equal(kf.weight.cpu().numpy().view(np.int32), want_w.astype(np.float32).view(np.int32))
    if kf.weight64 is not None:
        assert np.array_equal(kf.weight64.cpu().numpy(), want_w)


@pytest.mark.parametrize("k,s,thr,topk,impl", [
    (6, 4, 0.0, 0.01, "auto"),        # D = 256: CTA-pair tensor kernel, histogram in its epilogue
    (6, 5, 0.0, 0.001, "auto"),       # D = 1024
    (6, 4, 0.3, 0.002, "tcgen05x2"),  # threshold and top-n together
    (6, 2, 0.0, 0.01, "dedup"),       # D = 16: duplicate-row classes, weighted class-pair histogram
    (7, 2, 0.5, 0.003, "dedup"),
    (6, 2, 0.0, 0.01, "tcgen05"),     # narrow tensor kernel: candidates placed, then selected on the packed words
    (6, 2, 0.0, 0.01, "dp4a"),
])
def test_top_n_before_emission_matches_oracle(k, s, thr, topk, impl):
    from genomic_gnn_b200 import core

    kc, counts_np = _counts(k, s, 30_000, seed=11 + k + s)
    want_ei, want_w = fast.kf_edges_exact(counts_np, thr, topk)
    n = counts_np.shape[0]
    assert want_ei.shape[1] == int(n * n * topk)                       # the budget binds in every case above
    whole = core.kf_edges(kc.counts, kc.sqnorm, thr, topk, sq_max=kc.m**2, impl=impl, row_sum=kc.m)
    _check(whole, want_ei, want_w)
    assert whole.cutoff is not None and abs(whole.cutoff - want_w.min()) < 1e-15 and whole.tie_kept < whole.tie_total
    # chunked emission: a few row blocks at a time, down to chunks smaller than one row block's share
    for cap in (150_000, 40_000, 7_000):
        part = core.kf_edges(kc.counts, kc.sqnorm, thr, topk, sq_max=kc.m**2, impl=impl, row_sum=kc.m, chunk_records=cap)
        _check(part, want_ei, want_w)
    # no promise about the row sums: the tensor kernel counts classes in the global table directly
    plain = core.kf_edges(kc.counts, kc.sqnorm, thr, topk, sq_max=kc.m**2, impl=impl, row_sum=0, chunk_records=60_000)
    _check(plain, want_ei, want_w)
    # "discard": same work, one recycled chunk buffer, counts only
    gone = core.kf_edges(kc.counts, kc.sqnorm, thr, topk, sq_max=kc.m**2, impl=impl, row_sum=kc.m, chunk_records=40_000,
                         sink="discard")
    assert gone.n_edges == want_ei.shape[1] and not gone.resident and gone.n_candidates == whole.n_candidates


def test_in_kernel_class_table_equals_the_table_of_the_placed_candidates():
    """gg_kf_emit's class_hist (epilogue counters) against gg_kf_class_hist_packed over the same candidates."""
    import ctypes as C

    from genomic_gnn_b200 import _lib, core
    from genomic_gnn_b200.core import ptr

    kc, counts_np = _counts(6, 4, 30_000, seed=3)
    sq_max = kc.m**2
    for thr, m in ((0.0, kc.m), (0.0, 0), (0.4, kc.m)):
        hist = torch.zeros(core.kf_hist_bins(sq_max), dtype=torch.int64, device="cuda")
        recs, found, rowcnt, params, scored = core.kf_emit(kc.counts, kc.sqnorm, thr, sq_max, impl="tcgen05x2",
                                                           rec_capacity=1 << 22, class_hist=hist, hist_m=m)
        staged = core.StagedRecords(recs, found, rowcnt, params, scored)
        ordered = torch.empty(found, dtype=torch.int64, device="cuda")
        staged.place(ptr(ordered))
        ref = torch.zeros_like(hist)
        core.check(_lib.load().gg_kf_class_hist_packed(ptr(ordered), found, ptr(kc.sqnorm), sq_max * sq_max, ptr(ref),
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "hist")
        assert int(hist.sum()) == found and torch.equal(hist, ref)
        s, t, dot, prod = fast.kf_candidates(counts_np, thr)
        want = np.bincount(dot * (sq_max * sq_max + 1) + prod, minlength=hist.numel())
        assert np.array_equal(hist.cpu().numpy(), want)
    # a staging buffer far too small for the candidates: the table is complete all the same
    hist = torch.zeros(core.kf_hist_bins(sq_max), dtype=torch.int64, device="cuda")
    _, found, *_ = core.kf_emit(kc.counts, kc.sqnorm, 0.0, sq_max, impl="tcgen05x2", rec_capacity=1000, class_hist=hist,
                                hist_m=kc.m, retry=False)
    assert found > 1000 and int(hist.sum()) == found


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("k,small_k,thr,topk,n_reads,impl", [
    (6, "2,4", 0.0, 0.01, 20_000, "auto"),      # narrow kernel (placed candidates) + tensor kernel (streamed)
    (6, "2", 0.0, 0.01, 20_000, "dedup"),       # duplicate-row classes, streamed
    (7, "2,5", 0.8, 0.01, 30_000, "auto"),      # threshold-bound: nothing to cut
])
def test_sharded_build_equals_single(world, k, small_k, thr, topk, n_reads, impl):
    """Reads sharded by rank (global positions), KF sharded by row block: every rank's gathered result, and the
    concatenation of the rank shards when the result stays sharded, equal the single-rank build bit for bit -- the
    cut-off class included (the spread rule picks by global member number)."""
    from genomic_gnn_b200 import core, dist as ggdist

    reads = fast.synthetic_reads(n_reads, 150, seed=5)
    reads[::97, 60] = ord("N")
    ref = core.build_graph(reads, k, small_k, edges_threshold=thr, edges_keep_top_k=topk, kf_impl=impl)

    def rank_fn(comm, gather):
        lo, hi = ggdist.shard_bounds(n_reads, comm.world, comm.rank)
        shard = core.upload_reads(reads[lo:hi], pos_base=lo * 150)
        return core.build_graph(shard, k, small_k, edges_threshold=thr, edges_keep_top_k=topk, kf_impl=impl, comm=comm,
                                kf_gather=gather, kf_chunk_records=50_000)

    for gather in (True, False):
        got = ThreadWorld(world).run(lambda comm: rank_fn(comm, gather))
        for g in got:
            assert torch.equal(g.db.node_codes, ref.db.node_codes) and torch.equal(g.db.edge_index, ref.db.edge_index)
            assert torch.equal(g.db.weight, ref.db.weight)
        for i, want in enumerate(ref.kf):
            if gather:
                for g in got:
                    assert g.kf[i].n_candidates == want.n_candidates and not g.kf[i].sharded
                    assert torch.equal(g.kf[i].edge_index, want.edge_index) and torch.equal(g.kf[i].weight, want.weight)
            else:
                assert all(g.kf[i].sharded for g in got)
                assert [int(g.kf[i].edge_index.shape[1]) for g in got] == got[0].kf[i].per_rank
                assert torch.equal(torch.cat([g.kf[i].edge_index for g in got], dim=1), want.edge_index)
                assert torch.equal(torch.cat([g.kf[i].weight for g in got]), want.weight)


def _class_table_torch(counts_t, sq_max):
    """Exact score-class table of every node pair s < t with a positive dot, by arithmetic that shares nothing with
    the kernels under test: duplicate rows collapsed by torch.unique, dots by a cuBLAS half-precision GEMM (exact:
    counts <= 15, dots <= 225), classes counted by sorting the keys.  Unordered node pairs = (ordered class pairs
    weighted by multiplicities - diagonal) / 2 + pairs inside a class."""
    p_max = sq_max * sq_max
    rows, mult = torch.unique(counts_t, dim=0, return_counts=True)
    sq = (rows.long() ** 2).sum(1)
    rh = rows.half()
    table = torch.zeros((sq_max + 1) * (p_max + 1), dtype=torch.float64, device=counts_t.device)
    step = max(1, (1 << 27) // rows.shape[0])
    for a0 in range(0, rows.shape[0], step):
        dot = (rh[a0:a0 + step] @ rh.T).long()
        key = dot * (p_max + 1) + sq[a0:a0 + step, None] * sq[None, :]
        w = (mult[a0:a0 + step, None] * mult[None, :]).double()
        ok = dot > 0
        keys, inv = torch.unique(key[ok], return_inverse=True)
        table.index_add_(0, keys, torch.bincount(inv, weights=w[ok], minlength=keys.numel()))
    table = table.long()
    self_key = sq * (p_max + 1) + sq * sq
    table.index_add_(0, self_key, -(mult * mult))
    table //= 2
    table.index_add_(0, self_key, mult * (mult - 1) // 2)
    return table.cpu().numpy()


def test_reduced_config5_k9_sharded_top_one_percent():
    """BASELINE config 5 in miniature: k = 9 (262,144 nodes, 3.4e10 pairs per set), threshold 0, top 1 %, small_k 2 and
    5, reads and KF row blocks sharded over 4 ranks, results left sharded.  The budget int(N^2 p) = 687,194,767 edges
    per set binds and the candidates (~1.2e9 at small_k=5, ~3e10 at small_k=2) are never materialised.  Checked: the
    candidate count, the cut-off class, the members of it that are kept and the per-rank edge counts against an exact
    class table computed independently (_class_table_torch) and the spread rule; per-edge properties on samples of
    what each rank emitted; and full equality with the oracle on a 4,096-node sub-problem of the same matrices."""
    from genomic_gnn_b200 import core, dist as ggdist

    k, n_reads, topk = 9, 600_000, 0.01
    reads = fast.synthetic_reads(n_reads, 150, seed=9)
    world = 4

    def rank_fn(comm):
        lo, hi = ggdist.shard_bounds(n_reads, comm.world, comm.rank)
        shard = core.upload_reads(reads[lo:hi], pos_base=lo * 150)
        return core.build_graph(shard, k, "2,5", edges_threshold=0.0, edges_keep_top_k=topk, comm=comm, kf_gather=False,
                                kf_sink="discard", kf_chunk_records=1 << 25, keep_features=False)

    got = ThreadWorld(world).run(rank_fn)
    n = got[0].db.num_nodes
    assert n == 4**9
    budget = int(n * n * topk)
    blocks = ggdist.kf_block_ranges((n + 1023) // 1024, world)
    for i, kc in enumerate(got[0].kf_counts):
        kfs = [g.kf[i] for g in got]
        sq_max = kc.m ** 2
        p_max = sq_max * sq_max
        assert all(kf.per_rank == kfs[0].per_rank and kf.sharded and not kf.resident for kf in kfs)
        assert sum(kfs[0].per_rank) == budget
        table = _class_table_torch(kc.counts, sq_max)
        assert int(table.sum()) == kfs[0].n_candidates
        action, tie_budget, cutoff, tie_total, _ = core.choose_cutoff_class(table, p_max, budget)
        assert kfs[0].tie_total == tie_total and kfs[0].tie_kept == tie_budget and abs(kfs[0].cutoff - cutoff) < 1e-15
        assert 0 < tie_budget < tie_total
        # an equal share of the cut-off class on every rank: no rank is left with only the edges above it
        above = int(table[action == 1].sum())
        shares = np.array(kfs[0].per_rank, dtype=np.float64)
        assert shares.min() > 0.8 * budget / world and shares.max() < 1.2 * budget / world and above < budget
        # what each rank emitted last (still in its recycled buffer): s < t, s in the rank's rows, exact cosine as
        # weight, score at or above the cut-off
        c = kc.counts.cpu().numpy().astype(np.int64)
        sqn = (c ** 2).sum(1)
        for r, kf in enumerate(kfs):
            ei, w = kf.edge_index.cpu().numpy(), kf.weight.cpu().numpy()
            assert ei.shape[1] > 0
            pick = np.random.default_rng(r).integers(0, ei.shape[1], size=2000)
            s_, t_ = ei[0, pick], ei[1, pick]
            assert np.all(s_ < t_) and np.all(s_ >= blocks[r][0] * 1024) and np.all(s_ < blocks[r][1] * 1024)
            dot = (c[s_] * c[t_]).sum(1)
            cos = dot / np.sqrt((sqn[s_] * sqn[t_]).astype(np.float64))
            assert np.array_equal(w[pick].view(np.int32), cos.astype(np.float32).view(np.int32))
            assert np.all(cos >= cutoff * (1 - 1e-12))
    # full equality on a sub-problem of the same matrices
    for kc in got[0].kf_counts:
        sub, sqn = kc.counts[:4096].contiguous(), kc.sqnorm[:4096].contiguous()
        want_ei, want_w = fast.kf_edges_exact(sub.cpu().numpy(), 0.0, topk)
        kf = core.kf_edges(sub, sqn, 0.0, topk, sq_max=kc.m**2, row_sum=kc.m, chunk_records=30_000,
                           impl="dedup" if kc.counts.shape[1] == 16 else "auto")
        _check(kf, want_ei, want_w)
