"""Import the UNMODIFIED reference from /root/reference behind stand-ins.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Only usable in the build container
(the GPU box has no /root/reference); used by oracle/make_golden.py to generate
the committed vectors and by tests that skip when the reference is absent.

Two third-party modules the reference imports are not installed here and are not
vendored in the reference tree:

* ``torch_geometric.nn.MetaLayer`` (requirements.txt:11, unpinned; used at
  models/GNN.py:24, 146, 215).  It holds no arithmetic: it gathers ``x[row]``,
  ``x[col]``, calls ``edge_model`` then ``node_model`` and returns
  ``(x, edge_attr, u)``.  The stand-in below follows that published behaviour.
* ``skimage`` (requirements.txt:12, unpinned; image_to_graph_superpixel.py:4-5).
  ``img_as_float`` of a uint8 image is ``img / 255`` in float64; ``slic`` is
  replaced by a function returning a caller-supplied label map, so that the rest
  of the reference's superpixel function (np.unique, means, binary_dilation
  adjacency) runs unmodified.  SLIC itself stays unpinned.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GNC_REFERENCE_ROOT", "/root/reference")

_slic_labels = {"next": None}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "GNN.py"))


def _install_standins():
    import numpy as np
    import torch

    if "torch_geometric" not in sys.modules:
        class MetaLayer(torch.nn.Module):
            def __init__(self, edge_model=None, node_model=None, global_model=None):
                super().__init__()
                self.edge_model = edge_model
                self.node_model = node_model
                self.global_model = global_model
                for m in (edge_model, node_model, global_model):
                    if m is not None and hasattr(m, "reset_parameters"):
                        m.reset_parameters()

            def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
                row, col = edge_index[0], edge_index[1]
                if self.edge_model is not None:
                    edge_attr = self.edge_model(x[row], x[col], edge_attr, u,
                                                batch if batch is None else batch[row])
                if self.node_model is not None:
                    x = self.node_model(x, edge_index, edge_attr, u, batch)
                if self.global_model is not None:
                    u = self.global_model(x, edge_index, edge_attr, u, batch)
                return x, edge_attr, u

        tg = types.ModuleType("torch_geometric")
        tg_nn = types.ModuleType("torch_geometric.nn")
        tg_nn.MetaLayer = MetaLayer
        tg.nn = tg_nn
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = tg_nn

    if "skimage" not in sys.modules:
        def slic(image, n_segments=100, compactness=10.0, start_label=1, **kw):
            labels = _slic_labels["next"]
            if labels is None:
                raise RuntimeError("stand-in slic: call set_next_slic_labels() first")
            assert labels.shape == image.shape[:2]
            return labels

        def img_as_float(image):
            image = np.asarray(image)
            if image.dtype == np.uint8:
                return image.astype(np.float64) / 255.0
            return image.astype(np.float64)

        sk = types.ModuleType("skimage")
        seg = types.ModuleType("skimage.segmentation")
        util = types.ModuleType("skimage.util")
        seg.slic = slic
        util.img_as_float = img_as_float
        sk.segmentation, sk.util = seg, util
        sys.modules["skimage"] = sk
        sys.modules["skimage.segmentation"] = seg
        sys.modules["skimage.util"] = util


def set_next_slic_labels(labels):
    """Label map the stand-in ``slic`` returns on its next calls."""
    _slic_labels["next"] = labels


def load_reference():
    """Returns a namespace with the reference's own modules (unmodified)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    _install_standins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    ns = types.SimpleNamespace()
    ns.GNN = importlib.import_module("models.GNN")
    ns.MLP = importlib.import_module("models.MLP")
    ns.optimized = importlib.import_module("utils.image_to_graph.image_to_graph_optimized")
    ns.patch = importlib.import_module("utils.image_to_graph.image_to_graph_patch")
    ns.superpixel = importlib.import_module("utils.image_to_graph.image_to_graph_superpixel")
    ns.train_model = importlib.import_module("utils.train_model")
    return ns
