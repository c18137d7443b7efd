"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (no kernels)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genomic_gnn_b200 import dist as ggdist
from genomic_gnn_b200.core import CountResult, choose_cutoff, top_n_budget
from oracle import fast


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _histogram_of_shard(reads, k, pos_base):
    """What K1 leaves on a rank, computed by the oracle: (k+1)-mer histogram and
    first occurrences with global stream offsets (fixed-length rows, no N)."""
    u, v, cnt, first = fast.count_pairs(reads, k)
    # oracle stream has one separator per read; the device stream for uint8 [R, L] input has none
    length = reads.shape[1]
    first = first - first // (length + 1) + pos_base
    bins = 4 ** (k + 1)
    hist = np.zeros(bins, dtype=np.int32)
    fp = np.full(bins, -1, dtype=np.int64)  # GG_NO_POS viewed as int64
    e = (u << 2) | (v & 3)
    hist[e] = cnt
    fp[e] = first
    return hist, fp


def _worker(rank, world, port, k, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = ggdist.TorchComm()
        assert comm.world == world and comm.rank == rank and comm.device.type == "cpu"
        assert comm.sum_int(rank + 1) == world * (world + 1) // 2
        assert comm.gather_ints(10 * rank) == [10 * r for r in range(world)]
        # all-gather(v): ragged edge lists concatenate in rank order
        mine = torch.arange(2 * (rank + 2), dtype=torch.int64).reshape(2, rank + 2) + 100 * rank
        cat = comm.gather_cat(mine, dim=1)
        want = torch.cat([torch.arange(2 * (r + 2), dtype=torch.int64).reshape(2, r + 2) + 100 * r for r in range(world)], dim=1)
        assert torch.equal(cat, want)
        assert comm.gather_cat(torch.zeros(0), dim=0).numel() == 0
        # packed all-gather(v): several ragged tensors in one collective
        sizes = [r + 2 for r in range(world)]
        w32 = torch.arange(rank + 2, dtype=torch.float32) + rank
        w64 = torch.arange(rank + 2, dtype=torch.float64) * 2 + rank
        ei, f64, f32 = comm.gather_packed([mine, w64, w32], sizes, [1, 0, 0])
        assert torch.equal(ei, want)
        assert torch.equal(f32, torch.cat([torch.arange(r + 2, dtype=torch.float32) + r for r in range(world)]))
        assert torch.equal(f64, torch.cat([torch.arange(r + 2, dtype=torch.float64) * 2 + r for r in range(world)]))
        z = comm.gather_packed([torch.zeros((2, 0), dtype=torch.int64)], [0] * world, [1])
        assert z[0].shape == (2, 0)
        # zero-copy exchange: every rank owns a slice of buffers sized for all ranks
        total, lo = sum(sizes), sum(sizes[:rank])
        a = torch.full((total,), -1, dtype=torch.int64)
        b = torch.full((total,), -1.0, dtype=torch.float32)
        a[lo : lo + sizes[rank]] = torch.arange(sizes[rank]) + 1000 * rank
        b[lo : lo + sizes[rank]] = rank
        for req in comm.exchange_slices([a, b], sizes):
            req.wait()
        assert torch.equal(a, torch.cat([torch.arange(sizes[r]) + 1000 * r for r in range(world)]))
        assert torch.equal(b, torch.cat([torch.full((sizes[r],), float(r)) for r in range(world)]))
        assert comm.exchange_slices([torch.zeros(0)], [0] * world) == []

        # C1: sharded counting merges to the single-process result
        reads = fast.synthetic_reads(600, 50, seed=4)
        lo, hi = ggdist.shard_bounds(reads.shape[0], world, rank)
        hist, fp = _histogram_of_shard(reads[lo:hi], k, pos_base=lo * reads.shape[1])
        cr = CountResult(k, torch.from_numpy(hist), torch.from_numpy(fp), torch.zeros((4, 4), dtype=torch.int32), 0,
                         torch.zeros(8, dtype=torch.int64))
        merged = ggdist.reduce_counts(cr, comm)
        g_hist, g_fp = _histogram_of_shard(reads, k, 0)
        assert np.array_equal(merged.hist.numpy(), g_hist)
        occupied = g_hist > 0
        assert np.array_equal(merged.first_pos.numpy()[occupied], g_fp[occupied])
        assert merged.n_nonzero == -1   # left unknown on purpose (saves a device round trip)

        # a scalar that already lives on the "device": gathered without an upload
        assert comm.gather_device_ints(torch.tensor([10 + rank, 7], dtype=torch.int64)) == [[10 + r, 7] for r in range(world)]

        # bridges: only rank 1 has some
        br = torch.tensor([[1, 2, 7, 0], [3, 4, 9, 0]], dtype=torch.int32)
        st = torch.zeros(8, dtype=torch.int64)
        st[1] = 2 if rank == 1 else 0                       # GG_ST_BRIDGES as K1 leaves it on the device
        cr2 = CountResult(k, torch.from_numpy(hist.copy()), torch.from_numpy(fp.copy()),
                          br if rank == 1 else torch.zeros((2, 4), dtype=torch.int32), -1, st)
        m2 = ggdist.reduce_counts(cr2, comm)
        assert m2.n_bridges == 2 and torch.equal(m2.bridges, br)
        # status words travel with the histogram: illegal bytes on any rank raise on every rank, and a bridge list
        # that outgrew its buffer asks for a re-count
        st_bad = torch.zeros(8, dtype=torch.int64)
        st_bad[0] = 3 if rank == 0 else 0
        try:
            ggdist.reduce_counts(CountResult(k, torch.from_numpy(hist.copy()), torch.from_numpy(fp.copy()),
                                             torch.zeros((2, 4), dtype=torch.int32), -1, st_bad), comm)
            raise AssertionError("expected ValueError")
        except ValueError:
            pass
        st_big = torch.zeros(8, dtype=torch.int64)
        st_big[1] = 5 if rank == 1 else 0
        try:
            ggdist.reduce_counts(CountResult(k, torch.from_numpy(hist.copy()), torch.from_numpy(fp.copy()),
                                             torch.zeros((2, 4), dtype=torch.int32), -1, st_big), comm)
            raise AssertionError("expected OverflowError")
        except OverflowError as exc:
            assert exc.args[0] == 5
        open(os.path.join(tmp, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 4, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds_cover():
    for n, w in [(10, 3), (1_000_000, 8), (5, 8), (0, 2)]:
        b = [ggdist.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_kf_block_ranges_balanced():
    for nb, w in [(64, 8), (64, 2), (1024, 8), (16, 4), (3, 8), (1, 1)]:
        r = ggdist.kf_block_ranges(nb, w)
        assert r[0][0] == 0 and r[-1][1] == nb and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        work = [sum(nb - i for i in range(a, b)) for a, b in r]
        assert sum(work) == nb * (nb + 1) // 2
    work = [sum(64 - i for i in range(a, b)) for a, b in ggdist.kf_block_ranges(64, 8)]
    assert max(work) <= 1.2 * sum(work) / 8


def test_choose_cutoff_matches_oracle_policy():
    """Host half of K5 against the oracle's statement of the same policy."""
    rng = np.random.default_rng(0)
    counts = fast.subkmer_counts(rng.permutation(4**4), 4, 2)      # 256 nodes, m = 3
    s, t, dot, prod = fast.kf_candidates(counts, 0.0)
    sq_max = 9
    p_max = sq_max * sq_max
    hist = np.bincount(dot * (p_max + 1) + prod, minlength=(sq_max + 1) * (p_max + 1)).astype(np.int64)
    for frac in (0.01, 0.05, 0.2):
        budget = top_n_budget(256, frac, s.shape[0])
        action, tie_budget, cutoff, tie_total = choose_cutoff(hist, p_max, budget)
        act = action[dot * (p_max + 1) + prod]
        keep = act == 1
        ties = np.nonzero(act == 2)[0]
        keep[ties[fast.spread_keep(len(ties), tie_budget)]] = True
        assert int(keep.sum()) == budget
        ei, w = fast.kf_edges_exact(counts, 0.0, frac)
        assert np.array_equal(ei, np.stack([s[keep], t[keep]]))
        if cutoff is not None:
            assert abs(w.min() - cutoff) < 1e-15 and tie_total == len(ties)


def test_top_n_budget_matches_reference_quirks():
    assert top_n_budget(4096, 0.01, 6_329_598) == 167_772
    assert top_n_budget(16384, 0.01, 1_206_888) == 1_206_888
    assert top_n_budget(5, 0.01, 9) == 9
    assert top_n_budget(100, None, 7) == 7


def test_knn_row_shards_tile_the_rows():
    """Row f4b shards rows by 128-row blocks: contiguous, aligned, disjoint, covering, balanced to one block."""
    from genomic_gnn_b200.core import KNN_ROW_BLOCK, knn_row_shard

    for n in (1, 127, 128, 129, 4096, 65536, 12345):
        for world in (1, 2, 3, 8):
            shards = [knn_row_shard(n, world, r) for r in range(world)]
            assert shards[0][0] == 0 and shards[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
            assert all(lo % KNN_ROW_BLOCK == 0 and lo <= hi for lo, hi in shards)
            blocks = [-(-(hi - lo) // KNN_ROW_BLOCK) for lo, hi in shards]
            assert max(blocks) - min(blocks) <= 1


def test_spread_rule_keeps_exactly_the_budget_and_ignores_sharding():
    """The cut-off class rule of gg_kf_select_packed (core.tie_ratio / ties_admitted) against the oracle's statement:
    exactly B of T members, and a rank / chunk that starts at global member g0 needs nothing but g0."""
    from genomic_gnn_b200.core import tie_ratio, ties_admitted

    rng = np.random.default_rng(5)
    cases = [(10, 3), (19416, 16476), (7, 6), (2, 1), (1000, 1), (12345, 6172), (2**20 + 3, 2**19)]
    cases += [(int(t), int(rng.integers(1, t))) for t in rng.integers(2, 50_000, size=20)]
    for total, budget in cases:
        mask = fast.spread_keep(total, budget)
        assert mask.shape == (total,) and int(mask.sum()) == budget
        r = tie_ratio(budget, total)
        assert r == fast.tie_ratio(budget, total) and 0 < r < 2**64
        prefix = np.concatenate([[0], np.cumsum(mask)])
        cuts = sorted({0, total, *map(int, rng.integers(0, total + 1, size=6))})
        for g in cuts:
            assert ties_admitted(g, r) == int(prefix[g])
        # evenly spread: any window holds its proportional share +- 1
        for a, b in zip(cuts[:-1], cuts[1:]):
            assert abs((prefix[b] - prefix[a]) - (b - a) * budget / total) <= 1.0
    # sizes of BASELINE config 5 (5.5e11 pairs): still exact in 64.64 fixed point
    total, budget = 17_203_456_789, 9_876_543_210
    r = tie_ratio(budget, total)
    assert ties_admitted(total, r) == budget and ties_admitted(0, r) == 0
    assert all(0 <= ties_admitted(g + 1, r) - ties_admitted(g, r) <= 1 for g in (0, 1, 12345, total - 2, total - 1))
    assert tie_ratio(0, 5) == 0 and tie_ratio(5, 5) == 0


def test_rational_tables_are_exact_for_strict_and_inclusive_tests():
    """dmin / fast_scale of the scoring kernels against brute force over every (dot, sq_s, sq_t): strict `>` for a
    threshold, `>=` for a cut-off class; the pre-filter never rejects a passing pair."""
    from fractions import Fraction

    from genomic_gnn_b200.core import ScoreTest, rational_tables, threshold_tables

    sq_max = 16
    tests = [ScoreTest.from_threshold(0.8), ScoreTest.from_threshold(0.0), ScoreTest.from_threshold(0.5),
             ScoreTest(1, 36, True), ScoreTest(4, 9, True), ScoreTest(9, 16, True), ScoreTest(1, 1, True),
             ScoreTest(4, 9, False)]
    for test in tests:
        dmin, scale, shift, p_max = rational_tables(test, sq_max)
        assert p_max == sq_max * sq_max and dmin.shape == (p_max + 1,)
        cut = Fraction(test.num, test.den)
        for si in range(1, sq_max + 1):
            for sj in range(1, sq_max + 1):
                p = si * sj
                for dot in range(0, sq_max + 1):
                    val = Fraction(dot * dot, p)
                    passes = val >= cut if test.inclusive else val > cut
                    if test.inclusive and test.num == 0:
                        continue
                    assert (dot >= int(dmin[p])) == passes, (test, dot, si, sj)
                    if passes:
                        assert ((dot * dot) << shift) > int(scale[si]) * sj, (test, dot, si, sj)
    d1, s1, sh1, _ = threshold_tables(0.8, sq_max)
    d2, s2, sh2, _ = rational_tables(ScoreTest.from_threshold(0.8), sq_max)
    assert np.array_equal(d1, d2) and np.array_equal(s1, s2) and sh1 == sh2


def test_choose_cutoff_class_reports_the_lowest_contributing_class():
    from fractions import Fraction

    from genomic_gnn_b200.core import choose_cutoff_class

    p_max = 16
    hist = np.zeros(5 * (p_max + 1), dtype=np.int64)
    hist[4 * (p_max + 1) + 16] = 10     # cos 1
    hist[2 * (p_max + 1) + 8] = 20      # cos^2 = 1/2
    hist[1 * (p_max + 1) + 4] = 30      # cos^2 = 1/4
    hist[2 * (p_max + 1) + 16] = 5      # cos^2 = 1/4 as well (same class, another (dot, P))
    action, tb, cutoff, tt, cut = choose_cutoff_class(hist, p_max, 40)
    assert (tb, tt, cut) == (10, 35, Fraction(1, 4)) and abs(cutoff - 0.5) < 1e-15
    assert action[1 * (p_max + 1) + 4] == 2 and action[2 * (p_max + 1) + 16] == 2 and action[2 * (p_max + 1) + 8] == 1
    action, tb, cutoff, tt, cut = choose_cutoff_class(hist, p_max, 30)   # lands exactly on a class boundary
    assert (tb, tt, cut) == (0, 0, Fraction(1, 2)) and cutoff is None and int((action == 2).sum()) == 0
