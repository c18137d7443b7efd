"""The C-ABI library builds, loads and exports every symbol include/gg_b200.h declares (CPU only: no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gg_b200.h")).read()
    return sorted(set(re.findall(r"GG_API\s+[\w\s\*]+?\b(gg_\w+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("gg_count_kp1mers", "gg_first_occurrence", "gg_db_build", "gg_kf_counts", "gg_kf_features_f32",
                 "gg_kf_emit", "gg_kf_finalize", "gg_kf_class_hist", "gg_kf_select"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__

    __graft_entry__.build()
    from genomic_gnn_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in gg_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == _declared()
    loaded = _lib.load()
    assert loaded.gg_abi_version() == 1
    assert loaded.gg_hist_bins(8) == 4**9
    assert loaded.gg_stream_padded_bytes(150) == 176


def test_argument_errors_are_reported_without_a_gpu():
    from genomic_gnn_b200 import _lib

    lib = _lib.load()
    assert lib.gg_count_reset(99, None, None, None, None) == -2        # GG_ERR_K
    assert b"k outside" in lib.gg_last_error()
    assert lib.gg_kf_emit(None, None, None, None, None, None, 0, None, None, None) == -1  # GG_ERR_ARG


def test_product_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import genomic_gnn_b200 as gg

    with pytest.raises(gg.GGError):
        gg.weighted_directed_edges(["ACGTACGT"], k=3, inlcudeUNK=False)
    with pytest.raises(gg.GGError):
        gg.cosine_similarity_to_edge_index_weight([[0.5, 0.5], [1.0, 0.0]])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "genomic_gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_swar_acgt_formula():
    """The exact all-ACGT short cut of K1 (csrc/gg_count.cu: acgt_low_bits / all_acgt16), restated on 32-bit words:
    a byte lane passes iff it is one of A, C, G, T -- for every byte value in every lane, whatever its neighbours."""
    import numpy as np

    rng = np.random.default_rng(0)
    m = 0xFFFFFFFF

    def low_bits(x):
        s1, s2, s4 = x >> 1, x >> 2, x >> 4
        g1 = x & (~s2 | s1)
        g2 = s2 & ~s1 & ~x
        return ((~s4 & g1) | (s4 & g2)) & m

    for lane in range(4):
        for b in range(256):
            for _ in range(8):
                w = [int(v) for v in rng.integers(0, 256, 4)]
                w[lane] = b
                x = w[0] | w[1] << 8 | w[2] << 16 | w[3] << 24
                fixed = ((x & 0xE8E8E8E8) ^ 0x40404040) & m
                ok = ((fixed >> (8 * lane)) & 0xFF) == 0 and ((low_bits(x) >> (8 * lane)) & 1) == 1
                assert ok == (b in b"ACGT"), (lane, b)
