"""Training pass of the width-128 core of GraphNet with a hand-scheduled backward.

One ``torch.autograd.Function`` covers everything between the encoders' thin first layers and the
decoder's last ``Linear(128, 1)``: encoder tails, the ``n_blocks`` edge / node processors
(models/GNN.py:57-64, 95-104, MetaLayer order :146, 215) and the decoder's two hidden layers
(models/GNN.py:289-295).  The forward is the kernel sequence of the op-by-op autograd path (``ops.tc_linear``
3xTF32, LayerNorm, ordered CSR aggregation), except that the three node-side products of a block (``P``, ``Q`` of the
edge processor's split first layer and the node processor's ``h``-side product) are ONE multi-mode launch of the
chained kernel (``h`` read once; fp16 two-piece operands, 4.7e-7 per product); what changes is the backward:

* gradients that autograd would add with separate elementwise passes (``h`` feeds three products and a
  residual, ``e`` a product and a residual, ``e'`` the aggregation and the next block) are accumulated
  in the data-gradient GEMM's epilogue (``addend``) or by the gather kernel (``accumulate``) -
  one extra row read instead of a read-read-write pass each;
* ReLU backward rides in the producing data-gradient's epilogue (``mask``) everywhere;
* weight gradients of the split first layers are written straight into column slices of the
  ``[128, 384]`` / ``[128, 256]`` gradient, biases come out of the same tcgen05 pass;
* activations are released as soon as their last consumer has run.

Same arithmetic as the op-by-op path, hence the same parity figures (logits 2.7e-6, gradients <= 8e-6 rel-L2 against
the CPU oracle; tests/test_gpu_tc_engine.py runs both).  ``backward`` releases the saved activations as it goes, so a
second backward through the same graph (``retain_graph=True``) is not supported.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import Tensor

from . import _lib, ops
from ._lib import check

_f32 = torch.float32
# Backward of a width-128 layer: "fused" = data + weight + bias gradient from one pass over (dZ, X)
# (csrc/tc_bwd.cu); "pair" = the round-1 schedule (tc_linear data gradient + tc_wgrad), kept as the implementation
# the fused kernel is tested against.
BWD = os.environ.get("GNC_BWD", "fused")
# Forward of a block's edge / node MLP: "chain" = ONE launch of the chained kernel that also writes what the backward
# reads (a1, a2, the LayerNorm input and its row statistics; fp16 two-piece operands like the inference path), "layer" =
# three per-layer launches (3xTF32) that write and re-read the hidden activations - the form "chain" is tested against.
FWD = os.environ.get("GNC_TRAIN_FWD", "chain")
# Backward of the x[row] / x[col] gathers: the edge gradient summed by source and by destination in one launch ("1") or two
AGG_PAIR = os.environ.get("GNC_AGG_PAIR", "1") != "0"
# Test hook (tests/test_gpu_tc_engine.py): when set to a dict, the forward records the sign pattern ``activation > 0``
# of every ReLU of the core under the reference's module path, e.g. ``("graph_processor.blocks.0.edge_model.
# edge_processor", 1)`` for ``model[1]`` - what a mask-conditioned gradient comparison against the oracle needs.
CAPTURE: Optional[dict] = None


def _capture(path: str, index: int, act: Tensor) -> None:
    if CAPTURE is not None:
        CAPTURE[(path, index)] = (act > 0)


def _bwd_layer(dZ: Tensor, X: Tensor, W: Tensor, *, mask: bool = False, addend: Optional[Tensor] = None,
               dW_out: Optional[Tensor] = None, want_db: bool = False, db_out: Optional[Tensor] = None,
               accumulate: bool = False):
    """``(dX, dW, db)`` of ``y = x @ W.T + b`` given ``dZ = dL/dy`` and the layer input ``X``; ``mask``: ``dX *= (X > 0)``
    (``X`` is a ReLU output, ``dX`` its pre-activation gradient); ``addend`` is added to ``dX``; ``dW_out``: a
    ``[128, 128]`` column slice the weight gradient is written to."""
    if BWD == "fused":
        return ops.tc_bwd_layer(dZ, X, W, mask=mask, addend=addend, dW_out=dW_out, want_db=want_db, db_out=db_out,
                                accumulate=accumulate)
    assert not accumulate and db_out is None, "in-place accumulation is a feature of the fused backward kernel"
    r = ops.tc_wgrad(dZ, X, out=dW_out, want_db=want_db)
    dW, db = r if want_db else (r, None)
    dX = ops.tc_linear(dZ, W, transpose_w=True, mask=X if mask else None, addend=addend)
    return dX, dW, db


def _ln_fwd(z: Tensor, gamma: Tensor, beta: Tensor, eps: float, res: Optional[Tensor]):
    M, D = z.shape
    y = torch.empty(M, D, dtype=_f32, device=z.device)
    mean = torch.empty(M, dtype=_f32, device=z.device)
    rstd = torch.empty(M, dtype=_f32, device=z.device)
    check(ops._call("layernorm_fwd", 0.0, 4.0 * M * D * (3 if res is not None else 2),
                    _lib.load().gnc_layernorm_fwd_f32, z.data_ptr(), ops._ld(z), M, D, gamma.data_ptr(),
                    beta.data_ptr(), float(eps), ops._p(res), ops._ld(res) if res is not None else 0, y.data_ptr(),
                    ops._ld(y), mean.data_ptr(), rstd.data_ptr(), ops._stream()), "layernorm_fwd")
    return y, mean, rstd


def _ln_bwd(dy: Tensor, z: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor, dg_out: Optional[Tensor] = None,
            db_out: Optional[Tensor] = None):
    """LayerNorm backward; with ``dg_out`` / ``db_out`` the affine gradients are ADDED to those tensors."""
    lib = _lib.load()
    M, D = z.shape
    dev = z.device
    dz = torch.empty(M, D, dtype=_f32, device=dev)
    acc = dg_out is not None
    dg = dg_out if acc else torch.empty(D, dtype=_f32, device=dev)
    db = db_out if acc else torch.empty(D, dtype=_f32, device=dev)
    ws_n = int(lib.gnc_layernorm_bwd_workspace(M, D))
    ws = ops._workspace(dev, ws_n)
    check(ops._call("layernorm_bwd", 0.0, 4.0 * M * D * 3, lib.gnc_layernorm_bwd_f32, dy.data_ptr(), ops._ld(dy),
                    z.data_ptr(), ops._ld(z), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), M, D, dz.data_ptr(),
                    ops._ld(dz), dg.data_ptr(), db.data_ptr(), int(acc), ws.data_ptr(), ws_n, ops._stream()), "layernorm_bwd")
    return dz, dg, db


def _relu_bwd(dY: Tensor, Y: Tensor):
    """dZ = dY * (Y > 0) and its column sums (bias gradient) in one pass."""
    lib = _lib.load()
    M, N = Y.shape
    dev = Y.device
    dZ = torch.empty(M, N, dtype=_f32, device=dev)
    db = torch.empty(N, dtype=_f32, device=dev)
    ws_n = int(lib.gnc_colsum_workspace(M, N))
    ws = ops._workspace(dev, ws_n)
    check(ops._call("relu_bwd_colsum", 0.0, 4.0 * M * N * 3, lib.gnc_relu_bwd_colsum_f32, dY.data_ptr(), ops._ld(dY),
                    Y.data_ptr(), ops._ld(Y), M, N, dZ.data_ptr(), ops._ld(dZ), db.data_ptr(), 0, ws.data_ptr(), ws_n,
                    ops._stream()), "relu_bwd_colsum")
    return dZ, db


def _tail_fwd(a1: Tensor, tail_params, eps_l: float, residual: Optional[Tensor]):
    """``LayerNorm(Linear(ReLU(Linear(a1)))) + residual`` behind a first layer's ReLU output ``a1``; returns the output and
    what ``_tail_bwd`` needs.  ``tail_params`` = (W2, b2, W4, b4, gamma, beta)."""
    W2, b2, W4, b4, g, bt = tail_params
    a2 = ops.tc_linear(a1, W2, bias=b2, relu=True)
    # last layer + LayerNorm + residual in one launch; the LayerNorm input and its row statistics are written out for
    # the backward by the same epilogue (no separate LayerNorm pass over z3)
    M = a2.shape[0]
    z3 = torch.empty(M, 128, dtype=_f32, device=a2.device)
    mean = torch.empty(M, dtype=_f32, device=a2.device)
    rstd = torch.empty(M, dtype=_f32, device=a2.device)
    y = ops.tc_linear(a2, W4, bias=b4, gamma=g, beta=bt, eps=eps_l, residual=residual, ln_save=(z3, mean, rstd))
    return y, (a1, a2, z3, mean, rstd)


def _mlp_fwd_chain(x: Tensor, W0c: Tensor, b0: Tensor, tail_params, eps_l: float, residual: Tensor, gather0, gather1=None):
    """``LayerNorm(W4 relu(W2 relu(x W0c^T + b0 + gathers) + b2) + b4) + residual`` in one launch; returns the output and
    the ``_tail_bwd`` tuple ``(a1, a2, z3, mean, rstd)``."""
    W2, b2, W4, b4, g, bt = tail_params
    M, dev = x.shape[0], x.device
    a1, a2, z3 = (torch.empty(M, 128, dtype=_f32, device=dev) for _ in range(3))
    mean, rstd = torch.empty(M, dtype=_f32, device=dev), torch.empty(M, dtype=_f32, device=dev)
    y = ops.tc_mlp_chain(x, [(W0c, b0), (W2, b2), (W4, b4)], gather0=gather0, gather1=gather1, gamma=g, beta=bt, eps=eps_l,
                         residual=residual, stash=(a1, a2, z3, mean, rstd))
    return y, (a1, a2, z3, mean, rstd)


def _grad_sinks(params) -> Optional[list]:
    """``[p.grad for p in params]`` when the step accumulates in place (``ops.ACCUMULATE_GRADS``, fused backward, every gradient
    already allocated as a contiguous fp32 tensor), else None: gradients are then returned to autograd."""
    if not ops.ACCUMULATE_GRADS or BWD != "fused":
        return None
    out = [ops.grad_sink(p) for p in params]
    return None if any(g is None for g in out) else out


def _tail_bwd(dy: Tensor, saved, tail_params, mask_a1: bool, sinks: Optional[list] = None):
    """Backward of ``_tail_fwd``: returns the gradient with respect to ``a1`` (times ``a1 > 0`` when ``mask_a1``: then
    it is the first layer's pre-activation gradient) and ``[dW2, db2, dW4, db4, dgamma, dbeta]`` (all None when the
    gradients were added into ``sinks``, the parameters' ``.grad`` tensors).  The residual's gradient is ``dy`` itself."""
    a1, a2, z3, mean, rstd = saved
    W2, _, W4, _, g, _ = tail_params
    acc = sinks is not None
    sk = sinks if acc else [None] * 6
    dz3, dg, dbt = _ln_bwd(dy, z3, mean, rstd, g, sk[4], sk[5])
    del z3
    dz2, dW4, db4 = _bwd_layer(dz3, a2, W4, mask=True, want_db=True, dW_out=sk[2], db_out=sk[3], accumulate=acc)
    del dz3, a2
    da1, dW2, db2 = _bwd_layer(dz2, a1, W2, mask=mask_a1, want_db=True, dW_out=sk[0], db_out=sk[1], accumulate=acc)
    return da1, ([None] * 6 if acc else [dW2, db2, dW4, db4, dg, dbt])


class MlpTailFn(torch.autograd.Function):
    """``_tail_fwd`` / ``_tail_bwd`` as an operator of its own (no residual): the edge encoder on the class table runs
    through exactly the kernels the per-edge encoder runs through inside ``GraphNetCoreFn``, so both give the same bits."""

    @staticmethod
    def forward(ctx, eps_l, a1, *tail_params):
        y, saved = _tail_fwd(ops._rows(a1), tail_params, eps_l, None)
        ctx.saved, ctx.tail_params = saved, tail_params
        return y

    @staticmethod
    def backward(ctx, dy):
        da1, grads = _tail_bwd(ops._rows(dy), ctx.saved, ctx.tail_params, False, _grad_sinks(ctx.tail_params))
        ctx.saved = None
        need = ctx.needs_input_grad
        return (None, da1 if need[1] else None, *[g if need[2 + i] else None for i, g in enumerate(grads)])


def core_param_list(gn) -> List[Tensor]:
    """Parameters the core consumes, in the order ``GraphNetCoreFn`` returns their gradients:
    per MLP tail ``W2 b2 W4 b4 gamma beta``; blocks add their first layers ``W0 b0`` / ``V0 c0`` in front;
    the decoder contributes ``W0 b0 W2 b2 W4 b4`` (``W4`` is its ``[1, 128]`` output layer)."""
    ps: List[Tensor] = []

    def tail(m):
        return [m[2].weight, m[2].bias, m[4].weight, m[4].bias, m[5].weight, m[5].bias]

    ps += tail(gn.node_encoder.model)
    ps += tail(gn.edge_encoder.model)
    for blk in gn.graph_processor.blocks:
        em, nm = blk.edge_model.edge_processor.model, blk.node_model.node_processor.model
        ps += [em[0].weight, em[0].bias] + tail(em)
        ps += [nm[0].weight, nm[0].bias] + tail(nm)
    dec = gn.node_decoder.model
    ps += [dec[0].weight, dec[0].bias, dec[2].weight, dec[2].bias, dec[4].weight, dec[4].bias]
    return ps


class GraphNetCoreFn(torch.autograd.Function):
    """``(a1n [N,128], a1e [E,128]) -> y [N,1]``: ``a1n`` / ``a1e`` are the ReLU outputs of the encoders'
    first layers, ``y`` the decoder's output (its ``Linear(128, 1)`` is a row dot product whose backward also applies
    the ReLU mask of the layer below and yields that layer's pre-activation gradient in one pass).  ``eps`` = (node
    enc, edge enc, [edge proc, node proc] per block) LayerNorm epsilons."""

    @staticmethod
    def forward(ctx, graph, eps, n_blocks, edge_ready, a1n, a1e, *params):
        tcl = ops.tc_linear
        a1n, a1e = ops._rows(a1n), ops._rows(a1e)
        saved = []          # per MLP tail: (a1, a2, z3, mean, rstd)

        def tail(a1, p0, residual, eps_l, path):
            y, sv = _tail_fwd(a1, params[p0:p0 + 6], eps_l, residual)
            saved.append(sv)
            _capture(path, 1, sv[0])
            _capture(path, 3, sv[1])
            return y

        h = tail(a1n, 0, None, eps[0], "node_encoder")
        if edge_ready:      # ``a1e`` already is the encoded edge latent (class-table form, see GraphNet._forward_tc_train)
            e = a1e
            saved.append(None)
        else:
            e = tail(a1e, 6, None, eps[1], "edge_encoder")
        blocks = []
        for k in range(n_blocks):
            pe = 12 + 16 * k
            pn = pe + 8
            W0, b0, V0, c0 = params[pe], params[pe + 1], params[pn], params[pn + 1]
            # the two node-side products of the edge processor and the node processor's h-side product read the
            # same rows: one launch of the chained kernel in multi mode (h read once, kept in tensor memory)
            P, Q, T = ops.tc_linear_multi(h, [W0[:, 0:128], W0[:, 128:256], V0[:, 0:128]])
            e_in, h_in = e, h
            path_e = f"graph_processor.blocks.{k}.edge_model.edge_processor"
            path_n = f"graph_processor.blocks.{k}.node_model.node_processor"
            if FWD == "chain":
                e, sv = _mlp_fwd_chain(e_in, W0[:, 256:384], b0, params[pe + 2:pe + 8], eps[2 + 2 * k], e_in,
                                       (P, graph.src), (Q, graph.dst))
                del P, Q
                saved.append(sv)
                _capture(path_e, 1, sv[0])
                _capture(path_e, 3, sv[1])
                agg = ops._agg_raw(graph.dst_rowptr, graph.dst_eid, e, graph.num_nodes)
                h, sv = _mlp_fwd_chain(agg, V0[:, 128:256], c0, params[pn + 2:pn + 8], eps[3 + 2 * k], h_in, (T, None))
                del T
                saved.append(sv)
                _capture(path_n, 1, sv[0])
                _capture(path_n, 3, sv[1])
            else:
                a1 = tcl(e, W0[:, 256:384], bias=b0, relu=True, gather0=(P, graph.src), gather1=(Q, graph.dst))
                del P, Q
                e = tail(a1, pe + 2, e_in, eps[2 + 2 * k], path_e)
                agg = ops._agg_raw(graph.dst_rowptr, graph.dst_eid, e, graph.num_nodes)
                n1 = tcl(agg, V0[:, 128:256], bias=c0, relu=True, addend=T)
                del T
                h = tail(n1, pn + 2, h_in, eps[3 + 2 * k], path_n)
            blocks.append((h_in, e_in, agg))
        Wd0, bd0, Wd2, bd2, Wd4, bd4 = params[12 + 16 * n_blocks:12 + 16 * n_blocks + 6]
        d1 = tcl(h, Wd0, bias=bd0, relu=True)
        d2 = tcl(d1, Wd2, bias=bd2, relu=True)
        y = ops.dot_tail_fwd(d2, Wd4, bd4)
        _capture("node_decoder", 1, d1)
        _capture("node_decoder", 3, d2)
        ctx.graph, ctx.n_blocks, ctx.edge_ready = graph, n_blocks, edge_ready
        ctx.saved, ctx.blocks, ctx.dec = saved, blocks, (h, d1, d2)
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        # the ~33 fused backward-layer launches of the pass leave their per-CTA partial weight / bias gradients in
        # workspaces of their own; ONE launch reduces all of them when the pass is over (same order, same bits)
        with ops.DeferredBwdReduce():
            return GraphNetCoreFn._backward(ctx, dy)

    @staticmethod
    def _backward(ctx, dy):
        graph, n_blocks, params = ctx.graph, ctx.n_blocks, ctx.params
        saved, blocks = ctx.saved, ctx.blocks
        dev = dy.device
        n_tail = 6
        grads: List[Optional[Tensor]] = [None] * len(params)

        sinks = _grad_sinks(params)                 # the parameters' .grad tensors when accumulating in place
        acc = sinks is not None

        def sink(i, lo=None, hi=None):
            if not acc:
                return None
            return sinks[i] if lo is None else sinks[i][:, lo:hi]

        def tail_bwd(dy, idx, p0, mask_a1):
            sv = saved[idx]
            saved[idx] = None
            da1, g6 = _tail_bwd(dy, sv, params[p0:p0 + n_tail], mask_a1, sinks[p0:p0 + n_tail] if acc else None)
            grads[p0:p0 + n_tail] = g6
            return da1, sv[0]

        # parameter offsets
        p_node_enc, p_edge_enc = 0, n_tail
        p_blk = [2 * n_tail + k * 2 * (2 + n_tail) for k in range(n_blocks)]
        p_dec = 2 * n_tail + n_blocks * 2 * (2 + n_tail)

        # ---- decoder ----
        h_last, d1, d2 = ctx.dec
        Wd0, _, Wd2, _, Wd4, _ = params[p_dec:p_dec + 6]
        # Linear(128, 1) backward + the ReLU mask of d2 in one pass: dz2 = (dy * w4) * (d2 > 0), dW4 = dy^T d2, db4 = sum dy
        dz2, dWd4, dbd4 = ops.dot_tail_bwd(d2, Wd4, dy, relu_mask=True, dw_out=sink(p_dec + 4), db_out=sink(p_dec + 5))
        del d2
        dz1, dWd2, dbd2 = _bwd_layer(dz2, d1, Wd2, mask=True, want_db=True, dW_out=sink(p_dec + 2), db_out=sink(p_dec + 3),
                                     accumulate=acc)
        del dz2, d1
        dh, dWd0, dbd0 = _bwd_layer(dz1, h_last, Wd0, want_db=True, dW_out=sink(p_dec), db_out=sink(p_dec + 1), accumulate=acc)
        del dz1, h_last
        if not acc:
            grads[p_dec:p_dec + 6] = [dWd0, dbd0, dWd2, dbd2, dWd4.reshape(Wd4.shape), dbd4.reshape(params[p_dec + 5].shape)]
        ctx.dec = None

        # ---- blocks, last to first ----
        # Gradient sums.  The round-1 kernels add an incoming gradient in the data-gradient epilogue (``addend``).  The
        # fused kernel's epilogue owns one input feature per thread (32 rows each), so an addend would be read with
        # 4-byte loads queued behind the operand stream; there the partial products are written as they are and summed
        # by ONE streaming pass (gather_add_rows: up to four row sources) - the same number of row reads as an addend for
        # the edge latents (whose sum also takes the aggregation's gathered rows), two more [N, 128] rows for the nodes.
        in_kernel = BWD != "fused"
        de = None                                   # gradient with respect to the block's output e' (complete)
        de_part = None                              # its part that arrived through the NEXT block's edge MLP input
        for k in range(n_blocks - 1, -1, -1):
            h_in, e_in, agg = blocks[k]
            blocks[k] = None
            pe = p_blk[k]
            pn = pe + 2 + n_tail
            W0 = params[pe]
            V0 = params[pn]
            # node processor: h' = LN(MLP(cat[h, agg])) + h
            dn1, _ = tail_bwd(dh, 2 + 2 * k + 1, pn + 2, True)
            dV0 = sinks[pn] if acc else torch.empty(128, 256, dtype=_f32, device=dev)
            dagg, _, dc0 = _bwd_layer(dn1, agg, V0[:, 128:256], dW_out=dV0[:, 128:256], want_db=True, db_out=sink(pn + 1),
                                      accumulate=acc)
            del agg
            if in_kernel:
                dh, _, _ = _bwd_layer(dn1, h_in, V0[:, 0:128], addend=dh, dW_out=dV0[:, 0:128])   # + the residual's gradient
            else:
                t1, _, _ = _bwd_layer(dn1, h_in, V0[:, 0:128], dW_out=dV0[:, 0:128], accumulate=acc)
            if not acc:
                grads[pn], grads[pn + 1] = dV0, dc0
            del dn1
            # aggregation backward: every edge receives its destination's row, on top of what later blocks sent
            if de is None:
                de = ops._gather_raw(dagg, graph.dst)
            elif in_kernel:
                ops._gather_raw(dagg, graph.dst, out=de, accumulate=True)
            else:
                de = ops.gather_add_rows([de_part, de, dagg], [None, None, graph.dst])
                de_part = None
            del dagg
            # edge processor: e' = LN(MLP(cat[h[row], h[col], e])) + e
            da1, _ = tail_bwd(de, 2 + 2 * k, pe + 2, True)
            dW0 = sinks[pe] if acc else torch.empty(128, 384, dtype=_f32, device=dev)
            if AGG_PAIR:    # both sums from one pass over da1
                dP, dQ = ops._agg_pair_raw(graph.src_rowptr, graph.src_eid, graph.dst_rowptr, graph.dst_eid, da1, graph.num_nodes)
            else:
                dP = ops._agg_raw(graph.src_rowptr, graph.src_eid, da1, graph.num_nodes)
                dQ = ops._agg_raw(graph.dst_rowptr, graph.dst_eid, da1, graph.num_nodes)
            if in_kernel:
                de, _, db0 = _bwd_layer(da1, e_in, W0[:, 256:384], addend=de, dW_out=dW0[:, 256:384], want_db=True)   # + residual
            else:
                de_part, _, db0 = _bwd_layer(da1, e_in, W0[:, 256:384], dW_out=dW0[:, 256:384], want_db=True,
                                             db_out=sink(pe + 1), accumulate=acc)
            del da1, e_in
            if in_kernel:
                dh, _, _ = _bwd_layer(dP, h_in, W0[:, 0:128], addend=dh, dW_out=dW0[:, 0:128])
                del dP
                dh, _, _ = _bwd_layer(dQ, h_in, W0[:, 128:256], addend=dh, dW_out=dW0[:, 128:256])
                del dQ, h_in
            else:
                t2, _, _ = _bwd_layer(dP, h_in, W0[:, 0:128], dW_out=dW0[:, 0:128], accumulate=acc)
                del dP
                t3, _, _ = _bwd_layer(dQ, h_in, W0[:, 128:256], dW_out=dW0[:, 128:256], accumulate=acc)
                del dQ, h_in
                dh = ops.gather_add_rows([t1, t2, t3, dh], [None, None, None, None])
                del t1, t2, t3
            if not acc:
                grads[pe], grads[pe + 1] = dW0, db0
        if de_part is not None:                     # gradient of the encoded edge latent: through block 0's MLP + its residual
            de = ops.gather_add_rows([de_part, de], [None, None])
            de_part = None
        # ---- encoder tails: the thin first layers mask their own ReLU ----
        if ctx.edge_ready:
            da1e = de                               # gradient of the encoded edge latent itself
        else:
            da1e, _ = tail_bwd(de, 1, p_edge_enc, False)
        del de
        da1n, _ = tail_bwd(dh, 0, p_node_enc, False)
        ctx.saved = ctx.blocks = None
        need = ctx.needs_input_grad
        out_params = [g if need[6 + i] else None for i, g in enumerate(grads)]
        return (None, None, None, None, da1n if need[4] else None, da1e if need[5] else None, *out_params)


class ExpandClassRowsFn(torch.autograd.Function):
    """``table[idx]`` for a table of a few rows (edge classes of a grid) expanded to every edge; the backward is the
    per-class column sum of the incoming gradient, as two ordered CSR segmented sums (chunks of 512 edges, then the
    chunks of a class) - deterministic, no atomics."""

    @staticmethod
    def forward(ctx, table, idx, plan):
        ctx.plan, ctx.n_rows = plan, table.shape[0]
        return ops._gather_raw(ops._rows(table), idx)

    @staticmethod
    def backward(ctx, dout):
        rowptr1, eid1, rowptr2, eid2 = ctx.plan
        partial = ops._agg_raw(rowptr1, eid1, ops._rows(dout), int(rowptr1.shape[0]) - 1)
        return ops._agg_raw(rowptr2, eid2, partial, ctx.n_rows), None, None


def class_sum_plan(edge_class: Tensor, n_classes: int, chunk: int = 512):
    """CSRs for ``ExpandClassRowsFn.backward``: level 1 groups the edges of a class (ascending edge id) into chunks of
    ``chunk`` rows, level 2 groups the chunks of each class."""
    cls = edge_class.long()
    order = torch.argsort(cls, stable=True)
    counts = torch.bincount(cls, minlength=n_classes).cpu().tolist()
    bounds, seg_class, off = [0], [], 0
    for c, n in enumerate(counts):
        for lo in range(0, n, chunk):
            bounds.append(off + min(lo + chunk, n))
            seg_class.append(c)
        off += n
    dev = edge_class.device
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr1 = torch.tensor(bounds, **i32)
    n_seg = len(seg_class)
    per_class = [seg_class.count(c) for c in range(n_classes)]
    rp2 = [0]
    for n in per_class:
        rp2.append(rp2[-1] + n)
    return rowptr1, order.to(torch.int32), torch.tensor(rp2, **i32), torch.arange(max(n_seg, 1), **i32)[:n_seg]


def graphnet_core(gn, graph, a1n: Tensor, a1e: Tensor, edge_ready: bool = False) -> Tensor:
    """Run the width-128 core of ``gn`` (a tensor-core-eligible ``GraphNet``) with the hand-scheduled backward.
    ``edge_ready``: ``a1e`` is the encoded edge latent itself (the edge encoder ran elsewhere)."""
    eps = [gn.node_encoder.model[5].eps, gn.edge_encoder.model[5].eps]
    for blk in gn.graph_processor.blocks:
        eps += [blk.edge_model.edge_processor.model[5].eps, blk.node_model.node_processor.model[5].eps]
    n_blocks = len(gn.graph_processor.blocks)
    return GraphNetCoreFn.apply(graph, tuple(eps), n_blocks, bool(edge_ready), a1n, a1e, *core_param_list(gn))
