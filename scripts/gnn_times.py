"""Row f1 timings on the BASELINE config 4 graph (development aid): GCN 32_DB,32_KF0 forward + backward, SpMM GB/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from genomic_gnn_b200 import core, gnn
from genomic_gnn_b200.synth import synthetic_reads

built = core.build_graph(synthetic_reads(1_000_000, 150, seed=0), 8, "2,5", edges_threshold=0.8, edges_keep_top_k=0.01)
cfg = gnn.edges_config_from(built)
torch.manual_seed(0)
task = gnn.SamplingTask({"x": built.x[0]}, gnn.GCN(16, ["32_DB", "32_KF0"]), cfg, seed=0)
task.fit(3)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); task.fit(10); b.record(); torch.cuda.synchronize()
print(f"epoch (sample pairs + forward + backward + Adam): {a.elapsed_time(b) / 10:.3f} ms; pairs per epoch", sum(int(p[0].shape[1]) for p in task.pairs))
adj = task.encoder._adj["edge_index_KF0"]
x = torch.randn(built.db.num_nodes, 32, device="cuda")
for _ in range(3):
    gnn._spmm(adj.fwd, x)
a.record()
for _ in range(20):
    gnn._spmm(adj.fwd, x)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
nnz = int(adj.fwd[1].numel())
print(f"SpMM KF0: nnz {nnz}, d 32: {ms:.3f} ms = {nnz * (8 + 128) / ms / 1e6:.0f} GB/s of gather traffic (x is 8 MB: L2-resident)")
