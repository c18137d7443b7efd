"""Measured hardware ceilings for the two stages whose bound is not HBM bandwidth (bench.py's secondary rooflines).

    smem_atomic_rate()   shared-memory atomic adds per second, the update pattern of K1's bucket-count kernel
    tensor_i8_rate()     kind::i8 tcgen05 operations per second of the CTA-pair MMA stream K4 issues

Both launch a probe kernel of libgg_b200.so (gg_probe_*) and time it with CUDA events on the launching stream, in the
same process and at the same clocks as the numbers they are compared with."""
import ctypes as C

import torch

from . import _lib
from ._lib import GGError, check, ptr
from .core import _require_cuda, _stream


def _best_ms(launch, repeats):
    launch()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def smem_atomic_rate(bins: int = 32768, iters: int = 2048, repeats: int = 5) -> dict:
    _require_cuda()
    lib = _lib.load()
    scratch = torch.empty(1024, dtype=torch.int32, device="cuda")
    n = C.c_double(0.0)

    def launch():
        check(lib.gg_probe_smem_atomics(bins, iters, ptr(scratch), C.byref(n), _stream()), "gg_probe_smem_atomics")

    ms = _best_ms(launch, repeats)
    return {"updates_per_s": n.value / (ms * 1e-3), "ms": ms, "bins": bins, "updates": n.value}


def tensor_i8_rate(iters: int = 2000, n_kblocks: int = 8, repeats: int = 5) -> dict:
    _require_cuda()
    lib = _lib.load()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    n = C.c_double(0.0)

    def launch():
        check(lib.gg_probe_tensor_i8(iters, n_kblocks, ptr(err), C.byref(n), _stream()), "gg_probe_tensor_i8")

    ms = _best_ms(launch, repeats)
    if int(err.item()) != 0:
        raise GGError("tensor probe reported a barrier time-out")
    return {"tops": n.value / (ms * 1e-3) / 1e12, "ms": ms, "ops": n.value,
            "how": f"tcgen05.mma.cta_group::2.kind::i8 256x256x32, {n_kblocks} K blocks per tile, {iters} tiles per CTA "
                   f"pair, operands resident in shared memory, best of {repeats}"}
