"""``MLP`` with the reference's constructor, parameter names and forward contract
(reference models/MLP.py:5-47), computed by libgnc kernels.

The module tree (``self.model`` = Sequential of Linear / activation / norm) is kept
identical so checkpoints move both ways; ``forward`` walks that Sequential and maps
  Linear [+ ReLU]  -> ops.linear (bias and ReLU fused in the GEMM epilogue)
  LayerNorm        -> ops.layer_norm (residual fused when the caller passes one)
Activations other than ReLU and BatchNorm1d are applied with their own torch module
on the device (non-default configurations).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch.nn as nn
from torch import Tensor

from .. import ops


class MLP(nn.Module):
    def __init__(
        self,
        in_dim: int,
        out_dim: int,
        hidden_dim: int = 128,
        hidden_layers: int = 2,
        activation: str = "ReLU",
        initializer: None | str = None,
        norm_type: None | str = "LayerNorm",
    ):
        super().__init__()
        self.activation = getattr(nn, activation)()
        if initializer is not None:
            self.initializer = getattr(nn.init, initializer)
        # the first Linear + activation exist for every hidden_layers (0 included), as in the reference's layer list
        # (models/MLP.py:24-27): module indices, state_dict keys and parameter-creation order stay the same
        stack = [nn.Linear(in_dim, hidden_dim), self.activation]
        for _ in range(hidden_layers - 1):
            stack += [nn.Linear(hidden_dim, hidden_dim), self.activation]
        stack.append(nn.Linear(hidden_dim, out_dim))
        if norm_type is not None:
            assert norm_type in ["LayerNorm", "BatchNorm1d"]
            stack.append(getattr(nn, norm_type)(out_dim))
        self.model = nn.Sequential(*stack)
        if initializer is not None:
            for param in self.model.parameters():
                if param.requires_grad and len(param.shape) > 1:
                    self.initializer(param)

    # -- kernel path ------------------------------------------------------------
    def forward_segments(self, srcs: Sequence[Tensor], gathers: Optional[Sequence] = None,
                         residual: Optional[Tensor] = None) -> Tensor:
        """MLP over the column-concatenation of ``srcs`` (optionally row-gathered, see
        ops.linear) without materialising it; ``residual`` is added to the output."""
        mods = list(self.model)
        cur: Optional[Tensor] = None
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, nn.Linear):
                fuse_relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                if cur is None:
                    cur = ops.linear(srcs, m.weight, m.bias, relu=fuse_relu, gathers=gathers)
                else:
                    cur = ops.linear([cur], m.weight, m.bias, relu=fuse_relu)
                i += 2 if fuse_relu else 1
            elif isinstance(m, nn.LayerNorm) and m.elementwise_affine and m.bias is not None:
                last = i == len(mods) - 1
                cur = ops.layer_norm(cur, m.weight, m.bias, m.eps, residual if last else None)
                if last:
                    residual = None
                i += 1
            else:
                cur = m(cur)
                i += 1
        if residual is not None:
            cur = cur + residual
        return cur

    def forward(self, x: Tensor):
        x = x.reshape(x.size(0), -1)     # reference: x.view(x.size(0), -1); .float() happens in the op
        return self.forward_segments([x])
