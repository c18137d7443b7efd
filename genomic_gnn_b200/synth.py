"""Synthetic reads of the BASELINE configs: i.i.d. uniform bases,
``default_rng(seed).integers(0, 4, (R, L))`` through b"ACGT"."""
import numpy as np


def synthetic_reads(n_reads: int, length: int = 150, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=(n_reads, length), dtype=np.uint8)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
