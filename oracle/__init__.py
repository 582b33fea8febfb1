"""CPU oracle for the GraphNet_Classifier hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / CPU torch, what the reference computes on
the path named by BASELINE.json (image -> graph builders, GraphNet forward and
backward, classifier head).  It exists so that the CUDA product path under
``graphnet_classifier_b200/`` can be checked against it.

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
    ``--impl reference`` legs may import anything from ``oracle/``;
  * nothing under ``graphnet_classifier_b200/`` imports it - the product path fails
    loudly when the CUDA library is missing, it never falls back to this code.

Parity status (see DESIGN.md "Oracle"):
  * grid / pixel / patch builders and the label-map -> superpixel-graph stage:
    PINNED bit-exactly against the reference's own functions imported from
    /root/reference (``oracle/make_golden.py``), vectors committed under
    ``tests/golden/``;
  * GraphNet / CombinedModel forward, loss and every parameter gradient: PINNED
    against the unmodified reference ``models/GNN.py`` run behind a stand-in for
    the un-vendored ``torch_geometric.nn.MetaLayer`` (no arithmetic lives in it);
  * SLIC segmentation itself (scikit-image, un-vendored, unpinned version):
    PARITY UNPINNED - neither the reference nor this container ships it.
"""
