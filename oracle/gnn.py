"""Oracle restatement of the reference GraphNet / CombinedModel (CPU torch).

TEST INFRASTRUCTURE - see oracle/__init__.py.  This is an independent, compact
restatement of what ``models/GNN.py`` + ``models/MLP.py`` compute, with the module
tree laid out so that ``state_dict()`` keys, shapes and construction-time RNG
consumption are identical to the reference (checked by oracle/make_golden.py
against the real classes: same seed -> same tensors, and the shipped checkpoint
loads with every key matched).

It removes the dependency on ``torch_geometric`` (the reference only uses
``MetaLayer``, which does two index gathers and calls the edge then the node
model - models/GNN.py:146, 215) and on ``torch_scatter`` (the reference's own
fallback is ``index_add_`` - models/GNN.py:9-21).
"""
from __future__ import annotations

import torch
from torch import Tensor, nn


def scatter_sum(src: Tensor, index: Tensor, dim: int = 0, dim_size: int | None = None) -> Tensor:
    """models/GNN.py:9-21: rows of ``src`` summed into ``out[index]`` in edge order."""
    if dim != 0:
        raise NotImplementedError("scatter_sum supports dim=0 only")
    if src.ndim == 1:
        src = src[:, None]
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros(dim_size, src.shape[1], dtype=src.dtype, device=src.device)
    return out.index_add_(0, index.to(torch.long), src)


def _mlp_stack(in_dim, out_dim, hidden_dim, hidden_layers, activation, norm_type):
    """models/MLP.py:24-37: Linear/act pairs, final Linear, optional norm layer.

    Index layout for hidden_layers=2: 0 Linear, 1 act, 2 Linear, 3 act, 4 Linear,
    5 norm - which is where the checkpoint keys ``model.{0,2,4,5}`` come from.
    """
    act = getattr(nn, activation)()
    # the first Linear + activation exist for every hidden_layers, 0 included (models/MLP.py:24-25)
    seq = [nn.Linear(in_dim, hidden_dim), act]
    for _ in range(hidden_layers - 1):
        seq += [nn.Linear(hidden_dim, hidden_dim), act]
    seq.append(nn.Linear(hidden_dim, out_dim))
    if norm_type is not None:
        assert norm_type in ("LayerNorm", "BatchNorm1d")
        seq.append(getattr(nn, norm_type)(out_dim))
    return nn.Sequential(*seq)


class OracleMLP(nn.Module):
    """models/MLP.py:5-47."""

    def __init__(self, in_dim, out_dim, hidden_dim=128, hidden_layers=2, activation="ReLU",
                 initializer=None, norm_type="LayerNorm"):
        super().__init__()
        self.model = _mlp_stack(in_dim, out_dim, hidden_dim, hidden_layers, activation, norm_type)
        if initializer is not None:                       # models/MLP.py:39-43
            init = getattr(nn.init, initializer)
            for p in self.model.parameters():
                if p.requires_grad and p.dim() > 1:
                    init(p)

    def forward(self, x: Tensor) -> Tensor:
        return self.model(x.reshape(x.shape[0], -1).float())   # models/MLP.py:46-47


class _EdgeModel(nn.Module):
    """models/GNN.py:31-64."""

    def __init__(self, dn, de, hidden, layers, activation, initializer, norm_type):
        super().__init__()
        self.edge_processor = OracleMLP(2 * dn + de, de, hidden, layers, activation, initializer, norm_type)

    def forward(self, h_src, h_dst, e):
        return self.edge_processor(torch.cat([h_src, h_dst, e], dim=-1)) + e


class _NodeModel(nn.Module):
    """models/GNN.py:69-104."""

    def __init__(self, dn, de, hidden, layers, activation, initializer, norm_type):
        super().__init__()
        self.node_processor = OracleMLP(dn + de, dn, hidden, layers, activation, initializer, norm_type)

    def forward(self, h, edge_index, e):
        # dim_size=None in the reference (:99) resolves to index.max()+1, which is
        # N on every graph whose last node has an incoming edge; we follow the
        # reference literally so a mismatch surfaces as the same shape error.
        agg = scatter_sum(e, edge_index[1], dim=0)
        return self.node_processor(torch.cat([h, agg], dim=-1)) + h


class _MetaBlock(nn.Module):
    """What torch_geometric.nn.MetaLayer does with an edge and a node model and
    no global model (call sites models/GNN.py:146, 215)."""

    def __init__(self, edge_model, node_model):
        super().__init__()
        self.edge_model = edge_model
        self.node_model = node_model
        self.global_model = None

    def forward(self, h, edge_index, e):
        e = self.edge_model(h[edge_index[0]], h[edge_index[1]], e)
        h = self.node_model(h, edge_index, e)
        return h, e


class _Processor(nn.Module):
    """models/GNN.py:168-216."""

    def __init__(self, n_blocks, dn, de, hid_n, hid_e, lay_n, lay_e, activation, initializer, norm_type):
        super().__init__()
        self.blocks = nn.ModuleList()
        for _ in range(n_blocks):
            # edge model is constructed before the node model (models/GNN.py:146-165)
            em = _EdgeModel(dn, de, hid_e, lay_e, activation, initializer, norm_type)
            nm = _NodeModel(dn, de, hid_n, lay_n, activation, initializer, norm_type)
            self.blocks.append(_MetaBlock(em, nm))

    def forward(self, h, edge_index, e):
        for blk in self.blocks:
            h, e = blk(h, edge_index, e)
        return h, e


class OracleGraphNet(nn.Module):
    """models/GNN.py:222-309 (kwargs and defaults :230-254)."""

    def __init__(self, **kw):
        super().__init__()
        g = kw.get
        in_node = g("num_local_features", 3) + g("num_global_features", 0)
        in_edge = 1 + g("space_dim", 2)
        dn, de = g("out_dim_node", 128), g("out_dim_edge", 128)
        norm, act, init = g("norm_type", "LayerNorm"), g("activation", "ReLU"), g("initializer", None)
        self.name = "GraphNet"
        self.out_dim = g("out_channels", 1)
        self.node_encoder = OracleMLP(in_node, dn, g("hidden_dim_node", 128), g("hidden_layers_node", 2),
                                      act, init, norm)
        self.edge_encoder = OracleMLP(in_edge, de, g("hidden_dim_edge", 128), g("hidden_layers_edge", 2),
                                      act, init, norm)
        self.graph_processor = _Processor(
            g("n_blocks", 10), dn, de,
            g("hidden_dim_processor_node", 128), g("hidden_dim_processor_edge", 128),
            g("hidden_layers_processor_node", 2), g("hidden_layers_processor_edge", 2),
            act, init, norm)
        # decoder: no norm, default ReLU whatever `activation` says (models/GNN.py:289-295)
        self.node_decoder = OracleMLP(dn, self.out_dim, g("hidden_dim_decoder", 128),
                                      g("hidden_layers_decoder", 2), norm_type=None)

    @staticmethod
    def edge_geometry(pos: Tensor, edge_index: Tensor) -> Tensor:
        """models/GNN.py:299-302: [pos[dst]-pos[src], L1 distance]."""
        rel = pos[edge_index[1]] - pos[edge_index[0]]
        return torch.cat([rel, rel.abs().sum(dim=1, keepdim=True)], dim=1)

    def forward(self, x, pos, edge_index):
        e = self.edge_encoder(self.edge_geometry(pos, edge_index))
        h = self.node_encoder(x)
        h, _ = self.graph_processor(h, edge_index, e)
        return self.node_decoder(h)


class _Head(nn.Module):
    """models/GNN.py:312-325."""

    def __init__(self, in_features, classes):
        super().__init__()
        self.fc1 = nn.Linear(in_features, 128)
        self.fc2 = nn.Linear(128, 32)
        self.fc3 = nn.Linear(32, classes)
        self.relu = nn.ReLU()

    def forward(self, v):
        return self.fc3(self.relu(self.fc2(self.relu(self.fc1(v)))))


class OracleCombinedModel(nn.Module):
    """models/GNN.py:327-341, plus block-diagonal batching: when ``x`` holds
    ``B * num_nodes`` rows the decoder output is viewed ``[B, num_nodes * out_dim]``
    and logits come back ``[B, classes]`` (``[classes]`` for B == 1, as the
    reference returns)."""

    def __init__(self, graph_net=None, num_nodes=128 * 128, classes=2):
        super().__init__()
        self.graph_net = graph_net if graph_net is not None else OracleGraphNet()
        self.num_nodes = num_nodes
        self.classifier = _Head(num_nodes * self.graph_net.out_dim, classes)

    def forward(self, x, pos=None, edge_index=None):
        if pos is None and edge_index is None and isinstance(x, tuple):
            x, pos, edge_index = x
        y = self.graph_net(x, pos, edge_index)
        B = max(1, y.shape[0] // self.num_nodes)
        if B == 1:
            return self.classifier(y.flatten())
        return self.classifier(y.reshape(B, -1))


def padded_readout_logits(model: "OracleCombinedModel", graphs) -> Tensor:
    """Logits of graphs whose node count differs from ``model.num_nodes`` - OUR definition where the reference is
    undefined (models/GNN.py:339-340 flattens the ``[N, 1]`` node outputs into a head sized for exactly
    ``num_nodes`` values, main.py:65-66, and raises otherwise).  Per graph: GraphNet exactly as the reference runs it,
    then the first ``min(N, num_nodes)`` outputs, zero-padded to ``num_nodes``, through the reference's head.  Equal to
    ``model(graph)`` whenever ``N == num_nodes``."""
    rows = []
    for (x, pos, ei) in graphs:
        y = model.graph_net(x, pos, ei).flatten()
        v = torch.zeros(model.num_nodes, dtype=y.dtype)
        k = min(y.numel(), model.num_nodes)
        v[:k] = y[:k]
        rows.append(model.classifier(v))
    return torch.stack(rows)


def build_reference_config_model(resize_value: int, classes: int = 2, seed: int | None = 0,
                                 n_blocks: int = 3, num_nodes: int | None = None):
    """The model main.py:72-73 builds, seeded like BASELINE.md section 4."""
    if seed is not None:
        torch.manual_seed(seed)
    gn = OracleGraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=n_blocks)
    return OracleCombinedModel(gn, num_nodes=num_nodes or resize_value * resize_value, classes=classes)
