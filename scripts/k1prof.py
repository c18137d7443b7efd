import os, sys
sys.path.insert(0, "/root/repo")
import torch
from genomic_gnn_b200 import core
from genomic_gnn_b200.synth import synthetic_reads
reads = synthetic_reads(1_000_000, 150, seed=0)
stream = core.upload_reads(reads)
for k in (6, 7, 8):
    for _ in range(2):
        core.count_kp1mers(stream, k, impl="part", first_occurrence=False)
torch.cuda.synchronize()
