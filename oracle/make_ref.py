"""Stage the reference's own hot-path modules for the GPU box (test / measurement infrastructure only).

    python -m oracle.make_ref          # also run by __graft_entry__.build() when /root/reference is present

/root/reference does not exist on the GPU box, and the reference is pure Python, so there is nothing to compile:
this recipe places UNMODIFIED copies of the two modules that hold the hot path

    src/utils.py                                         seq_to_kmers, weighted_directed_edges, build_deBruijn_graph
    src/representations/gnn_common/gnn_utils.py          node_embedding_initial, cosine_similarity_to_edge_index_weight

under the git-ignored build directory ``oracle/_ref/`` (it travels with the gpurun snapshot like the compiled .so and
never enters the history), together with a manifest of their SHA-256 digests.  ``oracle/ref_import.py`` imports them
from there when /root/reference is absent, which lets ``bench.py --impl reference`` time the reference itself
(``cpu_baseline.kind == "reference"``) instead of the oracle's port.  Nothing under ``genomic_gnn_b200/`` reads it.
"""
import hashlib
import json
import os
import shutil

REFERENCE_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ["src/utils.py", "src/representations/gnn_common/gnn_utils.py"]
PACKAGES = ["src", "src/representations", "src/representations/gnn_common"]


def make() -> str:
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    manifest = {}
    for pkg in PACKAGES:
        os.makedirs(os.path.join(REF_DIR, pkg), exist_ok=True)
        init = os.path.join(REF_DIR, pkg, "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, rel), dst)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as fh:
        json.dump({"source": REFERENCE_ROOT, "sha256": manifest}, fh, indent=1)
    return REF_DIR


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, rel)) for rel in FILES)


if __name__ == "__main__":
    print(make())
