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
    assert loaded.gg_abi_version() == _lib.GG_ABI_VERSION == 3
    assert loaded.gg_hist_bins(8) == 4**9
    assert loaded.gg_stream_padded_bytes(150) == 176


def test_argument_errors_are_reported_without_a_gpu():
    from genomic_gnn_b200 import _lib

    lib = _lib.load()
    assert lib.gg_count_reset(99, None, None, None, None) == -2        # GG_ERR_K
    assert b"k outside" in lib.gg_last_error()
    assert lib.gg_kf_emit(None, None, None, None, None, None, 0, None, None, None, None) == -1  # GG_ERR_ARG


def test_newer_entry_points_validate_their_arguments_without_a_gpu():
    import ctypes as C

    from genomic_gnn_b200 import _lib

    lib = _lib.load()
    ERR_ARG, ERR_K = -1, -2
    assert lib.gg_count_part_workspace_bytes(150_000_000, 8) > 2 * 8 * 150_000_000   # 8 buckets of 16-bit keys
    assert lib.gg_count_part_workspace_bytes(150_000_000, 9) == 0                     # table too large: L2 path
    assert lib.gg_count_kp1mers_part(None, 16, 0, 9, 0, None, None, 0, None, None, 0, None) != 0
    assert lib.gg_count_merge(None, 2, 10, 8, None, None, None) == ERR_ARG
    assert lib.gg_count_merge(None, 2, 10, 99, None, None, None) == ERR_K
    assert lib.gg_kf_dedup_classes(None, 16, 64, 100, None, None, None, None, 0, None) == ERR_ARG    # > 16 columns
    assert lib.gg_kf_dedup_classes(None, 16, 16, (1 << 20) + 1, None, None, None, None, 0, None) == ERR_ARG
    assert lib.gg_kf_dedup_adj_words(17444) == 17444 * 546
    assert lib.gg_kf_dedup_score(None, 16, 16, None, None, None, 10, None, 50, None, None, None) == ERR_ARG  # p_max not a square
    assert lib.gg_kf_select_packed(None, 1 << 31, None, 16, None, 0, None, None, None, None, None, 0, None, 0, None) == ERR_ARG
    assert lib.gg_kf_select_packed(None, 10, None, 16, None, 0, None, None, None, None, None, 0, None, 0, None) == ERR_ARG
    assert lib.gg_kf_select_packed_workspace_bytes(1 << 20) >= 8 * (1 << 20)
    assert lib.gg_kf_class_hist_packed(None, 10, None, 16, None, None) == ERR_ARG
    assert lib.gg_kf_class_hist_packed(None, 0, None, 16, None, None) == 0
    assert lib.gg_probe_smem_atomics(1000, 1, None, None, None) == ERR_ARG                     # not a power of two
    assert lib.gg_probe_tensor_i8(0, 8, None, None, None) == ERR_ARG
    assert lib.gg_kf_dedup_count(100, 0, 5, None, None, None, 10, None, None, None, 0, None) == ERR_ARG    # shard past the end
    assert lib.gg_db_build(None, None, None, 0, 8, 1, 0, -1.0, None, None, None, None, None, None, 0, None, None, 0, 41,
                           None) == ERR_ARG
    assert lib.gg_rw_pairs(None, None, None, 10, None, 10, 1, 64, 3, 1.0, 1.0, 0, 0, None, None, None, None, 0, None,
                           None) == ERR_ARG                                                     # walk_length > 32
    assert b"walk_length" in lib.gg_last_error()
    assert lib.gg_rw_pairs(None, None, None, 10, None, 10, 1, 2, 3, 0.0, 1.0, 0, 0, None, None, None, None, 0, None,
                           None) == ERR_ARG                                                     # p must be positive
    one = C.c_void_p(1)   # never dereferenced: validation comes first
    assert lib.gg_knn_count(one, 24, 100, 0, 100, one, one, 49, one, 5, 49, 30, 0, one, None) == ERR_ARG    # pitch 24
    assert b"pitch" in lib.gg_last_error()
    assert lib.gg_knn_count(one, 16, 100, 0, 100, one, one, 49, one, 5, 49, 30, 3, one, None) == ERR_ARG    # hist_mode
    assert lib.gg_knn_count(one, 16, 100, 0, 100, one, one, 49, one, 5, 49, 2000, 2, one, None) == ERR_ARG  # does not fit
    assert lib.gg_knn_count(one, 16, 1000, 64, 500, one, one, 49, one, 5, 49, 30, 0, one, None) == ERR_ARG  # shard start
    assert lib.gg_knn_plan(100, 0, 100, 30, 100, 0, one, None, one, one, one, one, one, 0, one, 1 << 20, None) == ERR_ARG
    assert lib.gg_knn_plan(100, 0, 100, 30, 5, 1, one, None, one, one, one, one, one, 0, one, 1 << 20, None) == ERR_ARG
    assert lib.gg_knn_plan(100, 0, 100, 30, 5, 0, one, None, one, one, one, one, one, 3, one, 1 << 20, None) == ERR_ARG
    assert lib.gg_knn_plan_workspace_bytes(65536) >= 2 * 4 * 65537
    assert lib.gg_host_job_wait(None) == ERR_ARG
    assert not lib.gg_host_features_f32_start(None, 10, None, None, 4, 0, None)


def test_host_feature_expansion_is_the_table_lookup():
    """gg_host_features_f32 is host code: the CPU suite can run it.  Same 256-entry value table as the device kernel
    (core.value_lut), any thread count, sizes that do not divide evenly."""
    import ctypes as C

    import numpy as np

    from genomic_gnn_b200 import _lib
    from genomic_gnn_b200.core import value_lut

    lib = _lib.load()
    rng = np.random.default_rng(0)
    lut = value_lut(4).astype(np.float32)
    for n, threads in ((0, 1), (17, 1), (100_003, 3), (1 << 20, 8)):
        counts = rng.integers(0, 5, size=n, dtype=np.uint8)
        out = np.full(n, -1.0, dtype=np.float32)
        rc = lib.gg_host_features_f32(C.c_void_p(counts.ctypes.data), n, C.c_void_p(lut.ctypes.data),
                                      C.c_void_p(out.ctypes.data), threads)
        assert rc == 0 and np.array_equal(out, lut[counts])
    job = lib.gg_host_features_f32_start(C.c_void_p(counts.ctypes.data), n, C.c_void_p(lut.ctypes.data),
                                         C.c_void_p(out.ctypes.data), 4, 0, None)   # no CUDA event: starts at once
    assert job and lib.gg_host_job_wait(C.c_void_p(job)) == 0 and np.array_equal(out, lut[counts])


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


def test_round2_entry_points_validate_their_arguments_without_a_gpu():
    from genomic_gnn_b200 import _lib

    lib = _lib.load()
    ERR_ARG = -1
    assert lib.gg_kf_counts_upto(None, 16, None, 8, 2, None, None, None, None) == ERR_ARG      # the device row count
    assert lib.gg_knn_set_impl(4) == ERR_ARG
    assert lib.gg_knn_set_impl(1) == 0                                                        # mma.sync kernel: ...
    assert lib.gg_knn_write_scratch_words(1024, 65536, 5, 16, 48, 65536, 1000) == 1000        # ... one word per edge
    assert lib.gg_knn_set_impl(0) == 0
    assert lib.gg_knn_write_scratch_words(16, 65536, 40, 49, 700, 65536, 1000) == 1000        # narrow rows: never tcgen05
    assert lib.gg_knn_write(None, 1024, 4096, 0, 4096, None, None, 16, None, 5, 16, 48, 0, None, None, None, None, None,
                            None, 0, None, None, None, 0, None) == ERR_ARG
    # the sparse flavour: shapes it serves, index size, argument checks
    assert lib.gg_knn_sparse_eligible(1024, 65536, 5, 48) == 1
    assert lib.gg_knn_sparse_eligible(16, 65536, 40, 700) == 0            # narrow rows are dense
    assert lib.gg_knn_sparse_eligible(1024, 1 << 20, 5, 48) == 0          # one byte per node must fit shared memory
    assert lib.gg_knn_sparse_index_bytes(65536, 1024) > 65536 * 16 * 4
    assert lib.gg_knn_sparse_index(None, 1024, 65536, None, None, 16, 5, None, 0, None, None) == ERR_ARG
    assert lib.gg_knn_sparse_count(None, 0, 1024, 65536, 0, 65536, None, None, 5, 16, 48, None, None) == ERR_ARG
