"""N-GPU build == 1-GPU build (NCCL, one process per GPU).  Needs >= 2 GPUs; skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, k, small_k, thr, topk, n_reads):
    import torch.distributed as dist

    from genomic_gnn_b200 import core, dist as ggdist, hostshare
    from genomic_gnn_b200.synth import synthetic_reads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = ggdist.TorchComm()
        reads = synthetic_reads(n_reads, 150, seed=5)
        # sprinkle N so that bridge records exist on some ranks
        reads[::97, 60] = ord("N")
        lo, hi = ggdist.shard_bounds(n_reads, world, rank)
        shard = core.upload_reads(reads[lo:hi], pos_base=lo * 150)
        ref = core.build_graph(reads, k, small_k, edges_threshold=thr, edges_keep_top_k=topk) if rank == 0 else None

        def same(got):
            assert torch.equal(got.db.node_codes, ref.db.node_codes)
            assert torch.equal(got.db.edge_index, ref.db.edge_index)
            assert torch.equal(got.db.weight, ref.db.weight)
            for a, b in zip(got.kf, ref.kf):
                assert a.n_candidates == b.n_candidates
                assert torch.equal(a.edge_index, b.edge_index), (a.edge_index.shape, b.edge_index.shape)
                assert torch.equal(a.weight, b.weight)

        # 1. every exchange through NCCL
        got = core.build_graph(shard, k, small_k, edges_threshold=thr, edges_keep_top_k=topk, comm=comm)
        if rank == 0:
            same(got)
        # 2. count pieces, candidate counts and packed edges pushed over NVLink peer memory (three builds: the two
        #    halves of the symmetric buffer alternate)
        peer = comm.enable_peer_exchange(256 << 20)
        if peer:
            for _ in range(3):
                got = core.build_graph(shard, k, small_k, edges_threshold=thr, edges_keep_top_k=topk, comm=comm)
                if rank == 0:
                    same(got)
            comm.peers.check()
            assert comm.peers.used >= 6
        # 3. KF sets left sharded, every rank delivering its shard into the shared host arena
        sharded = core.build_graph(shard, k, small_k, edges_threshold=thr, edges_keep_top_k=topk, comm=comm, kf_gather=False)
        assert all(kf.sharded and int(kf.edge_index.shape[1]) == kf.per_rank[rank] for kf in sharded.kf)
        arena = hostshare.SharedHostArena(comm, hostshare.arena_bytes_for(
            sharded.db.num_nodes, k, small_k.split(","), [kf.n_edges for kf in sharded.kf]))
        host = hostshare.deliver_graph(sharded, arena, comm)
        if rank == 0:
            assert torch.equal(host.node_codes, ref.db.node_codes.cpu()) and torch.equal(host.edge_index, ref.db.edge_index.cpu())
            assert torch.equal(host.weight, ref.db.weight.cpu())
            for i, b in enumerate(ref.kf):
                assert torch.equal(host.x[i], ref.x[i].cpu())
                assert torch.equal(host.kf_edge_index[i], b.edge_index.cpu()) and torch.equal(host.kf_weight[i], b.weight.cpu())
            with open(os.path.join(out_dir, "ok"), "w") as fh:
                fh.write("peer" if peer else "nccl only: " + getattr(comm, "peer_error", "?"))
        del host
        arena.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,small_k,thr,topk,n_reads", [
    (7, "2,5", 0.8, 0.01, 40_000),     # threshold-bound, 16 row blocks
    (6, "2,4", 0.0, 0.01, 20_000),     # global top-n binding: class histogram all-reduce + tie split across ranks
    (8, "2", 0.8, 0.01, 120_000),      # 65,536 nodes: duplicate-row KF path sharded by row block, device-side count gather
    (8, "2", 0.8, 0.001, 120_000),     # ... with the global top-n binding (4.3M of 20.6M candidates kept)
])
def test_multi_gpu_equals_single(tmp_path, k, small_k, thr, topk, n_reads):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), k, small_k, thr, topk, n_reads), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")
    print("exchange:", open(tmp_path / "ok").read())


def _config5_worker(rank, world, port, out_dir):
    """BASELINE config 5 in miniature over NCCL: k = 9, threshold 0, top 1 %, both KF sets row-sharded and left
    sharded; rank 0 checks the concatenated shards of a 16-row-block prefix problem against the single-GPU result and
    the edge accounting of the full problem."""
    import torch.distributed as dist

    from genomic_gnn_b200 import core, dist as ggdist
    from genomic_gnn_b200.synth import synthetic_reads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = ggdist.TorchComm()
        comm.enable_peer_exchange(256 << 20)
        n_reads, k, topk = 400_000, 9, 0.01
        reads = synthetic_reads(n_reads, 150, seed=2)
        lo, hi = ggdist.shard_bounds(n_reads, world, rank)
        shard = core.upload_reads(reads[lo:hi], pos_base=lo * 150)
        got = core.build_graph(shard, k, "2,5", edges_threshold=0.0, edges_keep_top_k=topk, comm=comm, kf_gather=False,
                               kf_chunk_records=1 << 25, keep_features=False)
        n = got.db.num_nodes
        for kf in got.kf:
            assert kf.sharded and sum(kf.per_rank) == int(n * n * topk) and int(kf.edge_index.shape[1]) == kf.per_rank[rank]
            assert 0 < kf.tie_kept < kf.tie_total and min(kf.per_rank) > 0.8 * max(kf.per_rank)
            s_, t_ = kf.edge_index[0], kf.edge_index[1]
            assert bool((s_ < t_).all()) and bool((s_[1:] >= s_[:-1] - 1023).all())
        # a 16,384-node sub-matrix: sharded top-n, gathered, against the single-GPU run of the same call
        for kc in got.kf_counts:
            sub, sqn = kc.counts[:16384].contiguous(), kc.sqnorm[:16384].contiguous()
            b0, b1 = ggdist.kf_block_range(16, world, rank)
            impl = "dedup" if kc.counts.shape[1] == 16 else "auto"
            mine = core.kf_edges(sub, sqn, 0.0, topk, sq_max=kc.m**2, impl=impl, iblock_begin=b0, iblock_end=b1, comm=comm,
                                 row_sum=kc.m, chunk_records=400_000)
            if rank == 0:
                ref = core.kf_edges(sub, sqn, 0.0, topk, sq_max=kc.m**2, impl=impl, row_sum=kc.m)
                assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.weight, ref.weight)
        if rank == 0:
            open(os.path.join(out_dir, "ok"), "w").close()
    finally:
        dist.destroy_process_group()


def test_multi_gpu_reduced_config5(tmp_path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_config5_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")


def _knn_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from genomic_gnn_b200 import core, dist as ggdist
    from genomic_gnn_b200.synth import synthetic_reads

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = ggdist.TorchComm()
        built = core.build_graph(synthetic_reads(20_000, 150, seed=5), 7, "2,5", with_kf=False)   # replicated
        n = built.db.num_nodes
        for kc in built.kf_counts:
            for distance in ("IP", "L2"):
                got = core.kf_knn_edges(kc.counts, kc.sqnorm, int(0.01 * n), distance, comm=comm)
                if rank == 0:
                    ref = core.kf_knn_edges(kc.counts, kc.sqnorm, int(0.01 * n), distance)
                    assert torch.equal(got.edge_index, ref.edge_index) and torch.equal(got.weight, ref.weight)
        if rank == 0:
            open(os.path.join(out_dir, "ok"), "w").close()
    finally:
        dist.destroy_process_group()


def test_multi_gpu_knn_equals_single(tmp_path):
    """Row f4b: per-row kNN KF edges sharded by row range over the ranks (one integer exchanged for "L2", then an
    all-gather of the edge lists in rank order) == the single-GPU result."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    mp.spawn(_knn_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")
