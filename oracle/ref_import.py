"""Import shim for the reference's own hot-path functions.

/root/reference is read-only and does not exist on the GPU box.  In the dev container the functions are imported
from there (``oracle/make_golden.py``, CPU tests that are skipped when the tree is absent); on the GPU box from the
unmodified copies that ``oracle/make_ref.py`` staged under the git-ignored ``oracle/_ref/`` (the timed reference arm
of ``bench.py``).  It registers stub modules for the two imports
the reference makes at module top that are not installed here
(``torch_geometric.utils`` at gnn_utils.py:6 and ``faiss`` at gnn_utils.py:8);
neither is touched by the five hot-path functions.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def root():
    """Directory the reference modules are imported from: the tree itself, else the staged copies, else None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        return REFERENCE_ROOT
    if os.path.exists(os.path.join(STAGED_ROOT, "src", "representations", "gnn_common", "gnn_utils.py")):
        return STAGED_ROOT
    return None


def available() -> bool:
    return root() is not None


def load():
    """Return a namespace with the reference's five hot-path functions."""
    where = root()
    if where is None:
        raise RuntimeError("reference modules found neither at %s nor under %s" % (REFERENCE_ROOT, STAGED_ROOT))
    if where not in sys.path:
        sys.path.insert(0, where)
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgu = types.ModuleType("torch_geometric.utils")
        tgu.from_networkx = tgu.to_networkx = None
        tg.utils = tgu
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.utils"] = tgu
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))
    from src.utils import (  # utils.py:15,52,87
        seq_to_kmers,
        weighted_directed_edges,
        build_deBruijn_graph,
    )
    from src.representations.gnn_common.gnn_utils import (  # gnn_utils.py:240,428
        node_embedding_initial,
        cosine_similarity_to_edge_index_weight,
    )

    ns = types.SimpleNamespace(
        root=where,
        seq_to_kmers=seq_to_kmers,
        weighted_directed_edges=weighted_directed_edges,
        build_deBruijn_graph=build_deBruijn_graph,
        node_embedding_initial=node_embedding_initial,
        cosine_similarity_to_edge_index_weight=cosine_similarity_to_edge_index_weight,
    )
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints and shows tqdm bars inside the hot functions."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield
