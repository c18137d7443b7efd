"""CPU oracle for the k-mer graph-construction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``genomic_gnn_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the product.

Parity status
-------------
* ``faithful`` / ``fast`` restatements of ``src/utils.py`` and
  ``src/representations/gnn_common/gnn_utils.py``: PINNED against outputs of the
  reference's own functions, imported unmodified in the dev container by
  ``oracle/ref_import.py`` and frozen as ``tests/golden/*.npz`` by
  ``oracle/make_golden.py``.
* ``from_networkx`` (torch_geometric 2.3.1, third party, source not under
  /root/reference, no reference test pins it): PARITY UNPINNED -- the semantics
  are restated from the upstream behaviour; only the
  networkx ordering half could be verified here.
"""
