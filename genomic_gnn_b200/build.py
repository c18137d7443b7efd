"""Build libgg_b200.so (the C-ABI library of include/gg_b200.h) in-tree with nvcc for sm_100a.

    python -m genomic_gnn_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libgg_b200.so")
SOURCES = ["gg_common.cu", "gg_count.cu", "gg_db.cu", "gg_feat.cu", "gg_kf.cu", "gg_kf_mma.cu", "gg_kf_dedup.cu", "gg_index.cu", "gg_rw.cu", "gg_host.cu", "gg_spmm.cu", "gg_knn.cu", "gg_probe.cu", "gg_peer.cu"]
PYHOST_SRC = os.path.join(CSRC, "gg_pyhost.c")
PYHOST_LIB = os.path.join(HERE, "libgg_pyhost.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


BUILD_ID_MARK = b"GG_BUILD_ID="


def source_hash():
    """Digest of every file the library is built from: the build id compiled into the .so (gg_build_id())."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "gg_pyhost.c")
    files.append(os.path.join(os.path.dirname(HERE), "include", "gg_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:32]


def built_id(lib=LIB):
    """Build id embedded in an existing .so (None: absent or unreadable)."""
    try:
        with open(lib, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    at = blob.find(BUILD_ID_MARK)
    return None if at < 0 else blob[at + len(BUILD_ID_MARK): at + len(BUILD_ID_MARK) + 32].decode("ascii", "replace")


def is_current():
    return built_id() == source_hash()


def build_pyhost(force=False):
    """libgg_pyhost.so: the CPython-facing marshalling helper (plain C, gcc).  Rebuilt when its source is newer."""
    import sysconfig

    if not force and os.path.exists(PYHOST_LIB) and os.path.getmtime(PYHOST_LIB) >= os.path.getmtime(PYHOST_SRC):
        return PYHOST_LIB
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock_pyhost"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and os.path.exists(PYHOST_LIB) and os.path.getmtime(PYHOST_LIB) >= os.path.getmtime(PYHOST_SRC):
            return PYHOST_LIB
        tmp = PYHOST_LIB + ".tmp%d" % os.getpid()
        cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + sysconfig.get_paths()["include"],
               PYHOST_SRC, "-o", tmp, "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for gg_pyhost.c:\n%s\n%s" % (r.stdout, r.stderr))
        os.replace(tmp, PYHOST_LIB)
    return PYHOST_LIB


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f != "gg_pyhost.c"]
    files.append(os.path.join(os.path.dirname(HERE), "include", "gg_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def _compile(src):
    obj = os.path.join(BUILD, src.replace(".cu", ".o"))
    stamp = src == "gg_common.cu"   # carries the build id: always recompiled (1 s)
    if not stamp and os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime():
        return obj, ""
    extra = ['-DGG_BUILD_ID_STR="%s"' % source_hash()] if stamp else []
    cmd = [NVCC, *FLAGS, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile + link unless the .so already carries the build id of the current sources.  Safe to call from several
    processes at once (one rank per GPU): the work happens under an exclusive file lock and is re-checked inside."""
    os.makedirs(BUILD, exist_ok=True)
    build_pyhost(force)
    if not force and is_current():
        return LIB
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_current():
            return LIB
        return _build_locked(force, verbose)


def _build_locked(force, verbose):
    if force:
        for f in os.listdir(BUILD):
            if f != ".lock":
                os.remove(os.path.join(BUILD, f))
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(_compile, SOURCES))
    log = "\n".join(l for _, l in results if l)
    with open(os.path.join(BUILD, "ptxas.log"), "a") as fh:
        fh.write(log)
    if verbose:
        print(log)
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp, *[o for o, _ in results], "-Xcompiler", "-fPIC", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)   # atomic: a concurrent dlopen sees the old or the new file, never a partial one
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
