"""Row f2 of the scope table: sequences -> ``kmers_index_<k>`` arrays on the GPU.

    KmerSsgnn.transform / KmerSsgnnMb.transform      src/representations/kmer_ssgnn.py:100-160,
                                                     src/representations/kmer_ssgnn_miniBatch.py:98-158
    word2index                                       kmer_ssgnn.py:148-153  (seq_to_kmers(seq, k, stride, inlcudeUNK=True))
"""
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib, core
from ._lib import check, ptr
from .api import encode_kmers


def _index_lut(index_dict: Dict[str, int], k: int) -> Tuple[np.ndarray, int]:
    """code -> index table (-1 = k-mer not in index_dict) and the index of the all-N k-mer."""
    pad = "N" * k
    if pad not in index_dict:
        raise KeyError(pad)  # what the reference's word2index raises on the first N-holding window
    kmers = [km for km in index_dict if km != pad]
    lut = np.full(4**k, -1, dtype=np.int32)
    if kmers:
        lut[encode_kmers(kmers, k)] = np.fromiter((index_dict[km] for km in kmers), dtype=np.int64, count=len(kmers))
    return lut, int(index_dict[pad])


def kmers_to_index(sequences, index_dict: Dict[str, int], k: int, stride: int = 1):
    """word2index for a whole column of sequences.  Returns (list of int64 arrays, one per sequence;
    number of windows whose k-mer is missing from index_dict).  Missing windows hold -1: the caller decides
    (the reference rebuilds the graph before indexing, kmer_ssgnn.py:127-146)."""
    core._require_cuda()
    lib = _lib.load()
    seqs = list(sequences)
    dev = torch.device("cuda", torch.cuda.current_device())
    lut, unk = _index_lut(index_dict, k)
    n = len(seqs)
    lens = np.fromiter(map(len, seqs), dtype=np.int64, count=n)
    starts = np.zeros(n, dtype=np.int64)
    if n:
        starts[1:] = np.cumsum(lens[:-1] + 1)          # one separator byte after every read
    n_win = np.where(lens >= k, (lens - k) // stride + 1, 0)
    out_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(n_win, out=out_off[1:])
    total = int(out_off[-1])
    stream = core.upload_reads(seqs, device=dev)
    out = torch.empty(max(total, 1), dtype=torch.int64, device=dev)
    status = torch.empty(2, dtype=torch.int64, device=dev)
    d = lambda a: torch.from_numpy(a).to(dev)
    starts_d, lens_d, off_d, lut_d = d(starts), d(lens), d(out_off), d(lut)
    check(lib.gg_kmer_index(ptr(stream.data), ptr(starts_d), ptr(lens_d), n, k, int(stride), ptr(lut_d), unk,
                            ptr(off_d), ptr(out), ptr(status), core._stream()), "gg_kmer_index")
    st = status.cpu()
    if int(st[0]) > 0:
        raise ValueError(f"{int(st[0])} windows hold a byte outside ACGTN")
    flat = out[:total].cpu().numpy()
    return [flat[a:b] for a, b in zip(out_off[:-1], out_off[1:])], int(st[1])


def missing_kmers(sequences, index_dict: Dict[str, int], k: int, stride: int = 1) -> int:
    """How many windows carry a k-mer that index_dict does not know (kmer_ssgnn.py:117-125 builds the set)."""
    return kmers_to_index(sequences, index_dict, k, stride)[1]


def transform(self, dfs_list: list, index_dict: dict, k: int, args: dict):
    """Drop-in for KmerSsgnn.transform / KmerSsgnnMb.transform: adds the column ``kmers_index_<k>`` to every
    dataframe; if some k-mer is unknown it extends the dataset and rebuilds graph + embeddings through the object's
    own create_graphs / initialise_task / compute_embedding_layer, exactly like kmer_ssgnn.py:127-146."""
    stride = args["inference_stride"]
    columns = [df["sequence"].tolist() for df in dfs_list]
    results = [kmers_to_index(col, index_dict, k, stride) for col in columns]
    if any(missing for _, missing in results):
        for col in columns:
            self.ssgnn_dataset.extend(col)
        saved = [tm.encoder for tm in self.task_models_list]
        self.create_graphs(k=k)
        self.initialise_task(k=k)
        for tm, enc in zip(self.task_models_list, saved):
            tm.encoder = enc
        self.compute_embedding_layer(k=k)
        self.new_layer = self.layer
        index_dict = self.index_dict
        results = [kmers_to_index(col, index_dict, k, stride) for col in columns]
    for df, (arrays, _) in zip(dfs_list, results):
        df["kmers_index_" + str(k)] = arrays
    return dfs_list
