"""B200-native k-mer graph construction behind the genomic-gnn builder API.

Importing the package does not touch the GPU; every compute entry point needs a
CUDA device and libgg_b200.so (no CPU fallback).
"""
from .api import (  # noqa: F401
    GGError,
    ItemFrequency,
    KFLabels,
    KmerGraph,
    PairFrequency,
    build_deBruijn_graph,
    cosine_similarity_to_edge_index_weight,
    kf_faiss_to_edge_index_weight,
    node_embedding_initial,
    weighted_directed_edges,
)
from .builders import (  # noqa: F401
    create_graphs,
    create_graphs_mini_batch,
    kf_edges_config,
    networkx_to_dataloader_full_batch,
    networkx_to_dataloader_mini_batch,
)
from .core import build_graph, build_graph_to_host, graph_to_host  # noqa: F401
from . import gnn  # noqa: F401  (row f1: GCN encoder, contrastive trainer, edit-distance regressor)
from .sampler import sample_biased_random_walks, sample_biased_random_walks_miniBatch  # noqa: F401
from .transform import kmers_to_index, transform  # noqa: F401

__version__ = "0.1.0"
