"""Row f2: GPU word2index / transform vs the reference's per-read Python loop (restated by the oracle)."""
import types

import numpy as np
import pandas as pd
import pytest

from oracle import faithful, fast

pytestmark = pytest.mark.gpu


def _reads(n, seed, p_n=0.02):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ln = int(rng.integers(0, 90))
        s = "".join("ACGT"[c] for c in rng.integers(0, 4, ln))
        out.append("".join(("N" if rng.random() < p_n else ch) for ch in s) + "N" * int(rng.integers(0, 5)))
    return out


def _word2index(seq, index_dict, k, stride):
    # kmer_ssgnn.py:148-153
    return np.array([index_dict[km] for km in faithful.seq_to_kmers(seq, k, stride, inlcudeUNK=True)], dtype=np.int64)


@pytest.mark.parametrize("k,stride", [(3, 1), (4, 2), (6, 1), (5, 3)])
def test_kmers_to_index_matches_reference_loop(k, stride):
    import genomic_gnn_b200 as gg

    reads = _reads(400, seed=k * 10 + stride)
    kmers = ["".join(t) for t in __import__("itertools").product("ACGT", repeat=k)]
    rng = np.random.default_rng(1)
    perm = rng.permutation(len(kmers))
    index_dict = {"N" * k: len(kmers)}
    index_dict.update({km: int(i) for km, i in zip(kmers, perm)})
    got, missing = gg.kmers_to_index(reads, index_dict, k, stride)
    assert missing == 0 and len(got) == len(reads)
    for seq, arr in zip(reads, got):
        want = _word2index(seq, index_dict, k, stride)
        assert arr.dtype == np.int64 and np.array_equal(arr, want)


def test_missing_kmers_are_reported_and_bad_bytes_rejected():
    import genomic_gnn_b200 as gg

    index_dict = {"NNN": 0, "ACG": 1, "CGT": 2}
    got, missing = gg.kmers_to_index(["ACGT", "ACGA", "NN"], index_dict, 3, 1)
    assert missing == 1 and got[0].tolist() == [1, 2] and got[1].tolist() == [1, -1] and got[2].size == 0
    with pytest.raises(ValueError):
        gg.kmers_to_index(["ACGX"], index_dict, 3, 1)
    with pytest.raises(KeyError):
        gg.kmers_to_index(["ACGT"], {"ACG": 0}, 3, 1)


def test_transform_drop_in_adds_column_and_rebuilds_on_missing():
    import genomic_gnn_b200 as gg

    k = 3
    df = pd.DataFrame({"sequence": ["ACGTAC", "TTTNAC", "AC"]})
    calls = []
    full = {"NNN": 64}
    full.update({"".join(t): i for i, t in enumerate(__import__("itertools").product("ACGT", repeat=k))})

    def _rebuild(self_, name):
        calls.append(name)
        if name == "compute_embedding_layer":
            self_.index_dict = full
            self_.layer = "new-layer"

    me = types.SimpleNamespace(ssgnn_dataset=[], task_models_list=[types.SimpleNamespace(encoder="enc0")])
    me.create_graphs = lambda k: _rebuild(me, "create_graphs")
    me.initialise_task = lambda k: _rebuild(me, "initialise_task")
    me.compute_embedding_layer = lambda k: _rebuild(me, "compute_embedding_layer")
    partial = {"NNN": 64, "ACG": 6}
    out = gg.transform(me, [df], partial, k, {"inference_stride": 1})
    assert calls == ["create_graphs", "initialise_task", "compute_embedding_layer"]       # kmer_ssgnn.py:127-146
    assert me.ssgnn_dataset == df["sequence"].tolist() and me.new_layer == "new-layer"
    assert me.task_models_list[0].encoder == "enc0"
    col = out[0]["kmers_index_3"]
    for seq, arr in zip(df["sequence"], col):
        assert np.array_equal(arr, _word2index(seq, full, k, 1))
    # nothing missing -> no rebuild
    calls.clear()
    gg.transform(me, [df], full, k, {"inference_stride": 2})
    assert calls == []
    assert np.array_equal(df["kmers_index_3"][0], _word2index("ACGTAC", full, k, 2))
