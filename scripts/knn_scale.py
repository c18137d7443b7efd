"""Row f4b on N GPUs (torchrun): exact kNN KF edges sharded by row range; max-over-ranks device time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from genomic_gnn_b200 import core, dist as ggdist
from genomic_gnn_b200.synth import synthetic_reads

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
comm = ggdist.TorchComm()
built = core.build_graph(synthetic_reads(200_000, 150, seed=0), 8, "2,5", with_kf=False)
n = built.db.num_nodes
for kc in built.kf_counts:
    for gather in (False, True):
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = core.kf_knn_edges(kc.counts, kc.sqnorm, int(0.01 * n), "IP", comm=comm, gather=gather)
            b.record(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"world {world} D {kc.counts.shape[1]} gather={gather}: {t.item():.2f} ms, edges here {res.edge_index.shape[1]}", flush=True)
dist.destroy_process_group()
