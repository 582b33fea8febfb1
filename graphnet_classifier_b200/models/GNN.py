"""GraphNet / CombinedModel with the reference's public surface (reference
models/GNN.py), running on libgnc kernels.

Same class names, constructor kwargs, forward signatures, attribute names and
``state_dict`` keys as the reference (76 tensors for the shipped configuration:
``graph_net.{node,edge}_encoder.model.*``, ``graph_net.graph_processor.blocks.k.
{edge_model.edge_processor,node_model.node_processor}.model.*``,
``graph_net.node_decoder.model.*``, ``classifier.fc{1,2,3}.*``), so checkpoints load
both ways.  What differs is underneath:

* the ``x[row]`` / ``x[col]`` gathers and both ``torch.cat`` calls are never
  materialised - the first Linear of each processor reads its operand through
  gathered column segments (include/gnc.h, ``gnc_seg_t``);
* ``scatter_sum`` is an ordered CSR segmented sum (bit-identical to the CPU
  reference's ``index_add_`` order);
* ``torch_geometric.nn.MetaLayer`` (un-vendored, no arithmetic) is replaced by the
  small ``MetaLayer`` below with the same attribute names;
* block-diagonal batches are accepted: ``x`` may hold ``B * num_nodes`` rows, logits
  come back ``[B, classes]`` (``[classes]`` for one graph, as the reference returns).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from ..ops import GraphIndex, scatter_sum  # noqa: F401  (re-exported: reference module-level name)
from .MLP import MLP


class EdgeProcessor(nn.Module):
    """e' = MLP(cat[x_src, x_dst, e]) + e   (reference models/GNN.py:31-64)."""

    def __init__(self, in_dim_node: int, in_dim_edge: int, hidden_dim: int = 128, hidden_layers: int = 2,
                 activation: str = "ReLU", initializer: None | str = None, norm_type: None | str = "LayerNorm"):
        super().__init__()
        self.edge_processor = MLP(2 * in_dim_node + in_dim_edge, in_dim_edge, hidden_dim, hidden_layers,
                                  activation, initializer, norm_type)

    def forward(self, src, dest, edge_attr, u=None, batch=None):
        """MetaLayer calling convention: ``src`` / ``dest`` are already-gathered rows."""
        return self.edge_processor.forward_segments([src, dest, edge_attr], residual=edge_attr)

    def forward_graph(self, x: Tensor, graph: GraphIndex, edge_attr: Tensor) -> Tensor:
        """Same result from the un-gathered node latents: the gathers ride inside the
        first GEMM's operand loads; their backward is the CSR segmented sum."""
        gathers = [
            (graph.src, (graph.src_rowptr, graph.src_eid), graph.num_nodes),
            (graph.dst, (graph.dst_rowptr, graph.dst_eid), graph.num_nodes),
            None,
        ]
        return self.edge_processor.forward_segments([x, x, edge_attr], gathers=gathers, residual=edge_attr)


class NodeProcessor(nn.Module):
    """x' = MLP(cat[x, scatter_sum(e, col)]) + x   (reference models/GNN.py:69-104)."""

    def __init__(self, in_dim_node: int, in_dim_edge: int, hidden_dim: int = 128, hidden_layers: int = 2,
                 activation: str = "ReLU", initializer: None | str = None, norm_type: None | str = "LayerNorm"):
        super().__init__()
        self.node_processor = MLP(in_dim_node + in_dim_edge, in_dim_node, hidden_dim, hidden_layers,
                                  activation, initializer, norm_type)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Tensor, u=None, batch=None):
        return self.forward_graph(x, ops.graph_of(edge_index, x.size(0)), edge_attr)

    def forward_graph(self, x: Tensor, graph: GraphIndex, edge_attr: Tensor) -> Tensor:
        agg = ops.aggregate(edge_attr, graph)            # dim_size = N
        return self.node_processor.forward_segments([x, agg], residual=x)


class MetaLayer(nn.Module):
    """Stand-in for ``torch_geometric.nn.MetaLayer`` (reference models/GNN.py:24, 146):
    edge model on (x[row], x[col], edge_attr), then node model; no global model."""

    def __init__(self, edge_model=None, node_model=None, global_model=None):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = global_model

    def forward(self, x, edge_index, edge_attr=None, u=None, batch=None):
        graph = edge_index if isinstance(edge_index, GraphIndex) else ops.graph_of(edge_index, x.size(0))
        if self.edge_model is not None:
            if hasattr(self.edge_model, "forward_graph"):
                edge_attr = self.edge_model.forward_graph(x, graph, edge_attr)
            else:
                src = ops.gather_rows(x, graph.src, graph.src_rowptr, graph.src_eid)
                dst = ops.gather_rows(x, graph.dst, graph.dst_rowptr, graph.dst_eid)
                edge_attr = self.edge_model(src, dst, edge_attr, u, batch)
        if self.node_model is not None:
            if hasattr(self.node_model, "forward_graph"):
                x = self.node_model.forward_graph(x, graph, edge_attr)
            else:
                x = self.node_model(x, edge_index, edge_attr, u, batch)
        if self.global_model is not None:
            u = self.global_model(x, edge_index, edge_attr, u, batch)
        return x, edge_attr, u


def build_GN_block(in_dim_node: int, in_dim_edge: int, hidden_dim_node: int = 128, hidden_dim_edge: int = 128,
                   hidden_layers_node: int = 2, hidden_layers_edge: int = 2, activation: str = "ReLU",
                   initializer: None | str = None, norm_type: None | str = "LayerNorm"):
    """One message-passing block (reference models/GNN.py:110-165)."""
    edge_model = EdgeProcessor(in_dim_node, in_dim_edge, hidden_dim_edge, hidden_layers_edge, activation,
                               initializer, norm_type)
    node_model = NodeProcessor(in_dim_node, in_dim_edge, hidden_dim_node, hidden_layers_node, activation,
                               initializer, norm_type)
    return MetaLayer(edge_model=edge_model, node_model=node_model)


class GraphProcessor(nn.Module):
    """``n_iterations`` blocks applied in sequence (reference models/GNN.py:168-216)."""

    def __init__(self, n_iterations: int, in_dim_node: int, in_dim_edge: int, hidden_dim_node: int = 128,
                 hidden_dim_edge: int = 128, hidden_layers_node: int = 2, hidden_layers_edge: int = 2,
                 activation: str = "ReLU", initializer: None | str = None, norm_type="LayerNorm"):
        super().__init__()
        self.blocks = nn.ModuleList(
            build_GN_block(in_dim_node, in_dim_edge, hidden_dim_node, hidden_dim_edge, hidden_layers_node,
                           hidden_layers_edge, activation, initializer, norm_type)
            for _ in range(n_iterations))

    def forward(self, x, edge_index, edge_attr):
        for block in self.blocks:
            x, edge_attr, _ = block(x, edge_index, edge_attr)
        return x, edge_attr


_GRAPHNET_DEFAULTS = dict(
    num_global_features=0, num_local_features=3, space_dim=2, out_channels=1, n_blocks=10,
    out_dim_node=128, out_dim_edge=128,
    hidden_dim_node=128, hidden_dim_edge=128, hidden_dim_decoder=128,
    hidden_dim_processor_node=128, hidden_dim_processor_edge=128,
    hidden_layers_node=2, hidden_layers_edge=2, hidden_layers_decoder=2,
    hidden_layers_processor_node=2, hidden_layers_processor_edge=2,
    norm_type="LayerNorm", activation="ReLU", initializer=None,
)


class GraphNet(nn.Module):
    """Encode-process-decode GraphNet (reference models/GNN.py:222-309; kwargs and
    defaults :230-254)."""

    def __init__(self, **kwargs):
        super().__init__()
        c = dict(_GRAPHNET_DEFAULTS)
        c.update({k: v for k, v in kwargs.items() if k in c})
        in_dim_node = c["num_local_features"] + c["num_global_features"]
        in_dim_edge = 1 + c["space_dim"]
        self.name = "GraphNet"
        self.out_dim = c["out_channels"]
        common = dict(activation=c["activation"], initializer=c["initializer"], norm_type=c["norm_type"])
        self.node_encoder = MLP(in_dim_node, c["out_dim_node"], c["hidden_dim_node"], c["hidden_layers_node"], **common)
        self.edge_encoder = MLP(in_dim_edge, c["out_dim_edge"], c["hidden_dim_edge"], c["hidden_layers_edge"], **common)
        self.graph_processor = GraphProcessor(
            c["n_blocks"], c["out_dim_node"], c["out_dim_edge"],
            c["hidden_dim_processor_node"], c["hidden_dim_processor_edge"],
            c["hidden_layers_processor_node"], c["hidden_layers_processor_edge"], **common)
        # the decoder keeps the MLP defaults for activation and has no norm (reference :289-295)
        self.node_decoder = MLP(c["out_dim_node"], self.out_dim, c["hidden_dim_decoder"],
                                c["hidden_layers_decoder"], norm_type=None)

    def forward(self, x, pos, edge_index):
        graph = edge_index if isinstance(edge_index, GraphIndex) else ops.graph_of(edge_index, x.size(0))
        edge_attr = ops.edge_geometry(pos, graph)          # [pos[col]-pos[row], L1]  (reference :299-302)
        out = self.node_encoder(x)
        edge_attr = self.edge_encoder(edge_attr)
        out, _ = self.graph_processor(out, graph, edge_attr)
        return self.node_decoder(out)


class LinearClassifier(nn.Module):
    """fc1 -> ReLU -> fc2 -> ReLU -> fc3 (reference models/GNN.py:312-325)."""

    def __init__(self, in_features=128 * 128, classes=2):
        super().__init__()
        self.fc1 = nn.Linear(in_features=in_features, out_features=128)
        self.fc2 = nn.Linear(in_features=128, out_features=32)
        self.fc3 = nn.Linear(in_features=32, out_features=classes)
        self.relu = nn.ReLU()

    def forward(self, x):
        one = x.dim() == 1
        v = x.reshape(1, -1) if one else x
        v = ops.linear([v], self.fc1.weight, self.fc1.bias, relu=True)
        v = ops.linear([v], self.fc2.weight, self.fc2.bias, relu=True)
        v = ops.linear([v], self.fc3.weight, self.fc3.bias, relu=False)
        return v.reshape(-1) if one else v


class CombinedModel(nn.Module):
    """GraphNet + flatten readout + classifier head (reference models/GNN.py:327-341)."""

    def __init__(self, graph_net: Optional[GraphNet] = None, num_nodes: int = 128 * 128, classes: int = 2):
        super().__init__()
        self.graph_net = graph_net if graph_net is not None else GraphNet()
        self.num_nodes = num_nodes
        in_features = num_nodes * self.graph_net.out_dim
        self.classifier = LinearClassifier(in_features=in_features, classes=classes)

    def forward(self, x, pos=None, edge_index=None):
        if pos is None and edge_index is None and isinstance(x, tuple):
            x, pos, edge_index = x
        y = self.graph_net(x, pos, edge_index)
        n_graphs = y.shape[0] // self.num_nodes if y.shape[0] > self.num_nodes else 1
        if n_graphs <= 1:
            return self.classifier(y.flatten())            # [classes], as the reference
        return self.classifier(y.reshape(n_graphs, -1))    # [B, classes]
