"""Build libgg_b200.so (the C-ABI library of include/gg_b200.h) in-tree with nvcc for sm_100a.

    python -m genomic_gnn_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libgg_b200.so")
SOURCES = ["gg_common.cu", "gg_count.cu", "gg_db.cu", "gg_feat.cu", "gg_kf.cu", "gg_kf_mma.cu", "gg_kf_dedup.cu", "gg_index.cu", "gg_rw.cu", "gg_host.cu", "gg_spmm.cu", "gg_knn.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "gg_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def _compile(src):
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    if os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime():
        return obj, ""
    cmd = [NVCC, *FLAGS, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    log = "\n".join(l for _, l in results if l)
    with open(os.path.join(BUILD, "ptxas.log"), "a") as fh:
        fh.write(log)
    if verbose:
        print(log)
    cmd = [NVCC, "-shared", "-o", LIB, *[o for o, _ in results], "-Xcompiler", "-fPIC", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
