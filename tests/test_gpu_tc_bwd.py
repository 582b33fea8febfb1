"""Fused backward-layer kernel (csrc/tc_bwd.cu, gnc_tc_bwd_layer_f32) against float64: the data gradient, the weight
gradient and the bias gradient of ``y = x W^T + b`` (models/MLP.py:24-27 under autograd) from one pass over dZ and X.
Tolerance: 2e-6 rel-L2 per tensor (the kernel's fp16 two-piece products measure ~5e-7), masks exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def _case(M, seed, decades=0.0, addend=False, strided=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    dZ = torch.randn(M, 128, generator=g, device=dev)
    if decades:
        # rows of very different magnitude, as the gradients of different graphs of a batch have
        dZ = dZ * torch.pow(10.0, -decades * torch.rand(M, 1, generator=g, device=dev))
    X = torch.relu(torch.randn(M, 128, generator=g, device=dev) * 3.0)
    Wfull = (torch.rand(128, 384, generator=g, device=dev) - 0.5) * 0.18
    W = Wfull[:, 128:256] if strided else Wfull[:, :128].contiguous()
    ad = torch.randn(M, 128, generator=g, device=dev) * 1e-3 if addend else None
    return dZ, X, W, ad


def _check(M, seed, decades=0.0, mask=True, addend=False, strided=False, tol=2e-6):
    from graphnet_classifier_b200 import ops
    dZ, X, W, ad = _case(M, seed, decades, addend, strided)
    dX, dW, db = ops.tc_bwd_layer(dZ, X, W, mask=mask, addend=ad, want_db=True)
    torch.cuda.synchronize()
    ref = dZ.double() @ W.double()
    if mask:
        ref = ref * (X > 0)
    if ad is not None:
        ref = ref + ad.double()
    refW = dZ.double().t() @ X.double()
    refb = dZ.double().sum(0)
    assert torch.isfinite(dX).all() and torch.isfinite(dW).all()
    e = (_rel(dX, ref), _rel(dW, refW), _rel(db, refb))
    assert e[0] < tol and e[1] < tol and e[2] < tol, e
    if mask and ad is None:
        assert bool((dX[X <= 0] == 0).all())                # the ReLU mask is exact
    return e


@pytest.mark.parametrize("M", [1, 31, 32, 33, 128, 4096 + 17, 148 * 32 * 3 + 5])
def test_bwd_layer_shapes(M):
    _check(M, seed=M)


@pytest.mark.parametrize("mask,addend,strided", [(False, False, False), (True, True, False), (False, True, True), (True, False, True)])
def test_bwd_layer_epilogues(mask, addend, strided):
    _check(70001, seed=3, mask=mask, addend=addend, strided=strided)


def test_bwd_layer_gradient_dynamic_range():
    from graphnet_classifier_b200 import ops
    # 8 decades between rows: the per-block scaling keeps the sums (dW, db) and dX at full accuracy
    _check(200000, seed=5, decades=8.0)
    # uniformly tiny and uniformly huge gradients
    for scale in (1e-12, 1e-20, 1e6):
        dZ, X, W, _ = _case(50000, 7)
        dZ = dZ * scale
        dX, dW, db = ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True)
        ref = (dZ.double() @ W.double()) * (X > 0)
        assert _rel(dX, ref) < 2e-6 and _rel(dW, dZ.double().t() @ X.double()) < 2e-6, scale


def test_bwd_layer_large_activations():
    # activations far above the fixed-scale chained kernel's |a| < 4094 domain
    from graphnet_classifier_b200 import ops
    dZ, X, W, _ = _case(30000, 11)
    X = X * 1e4
    dX, dW, db = ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True)
    assert _rel(dW, dZ.double().t() @ X.double()) < 2e-6


def test_bwd_layer_accumulate_and_equals_separate_kernels():
    from graphnet_classifier_b200 import ops
    dZ, X, W, _ = _case(100003, 13)
    acc = torch.ones(128, 384, device="cuda")
    accb = torch.ones(128, device="cuda")
    ops.tc_bwd_layer(dZ, X, W, dW_out=acc[:, 128:256], accumulate=True, want_db=True, db_out=accb)
    refW = dZ.double().t() @ X.double()
    assert _rel(acc[:, 128:256] - 1, refW) < 2e-6 and bool((acc[:, :128] == 1).all()) and bool((acc[:, 256:] == 1).all())
    assert _rel(accb - 1, dZ.double().sum(0)) < 1e-5
    # against the round-1 pair of kernels (3xTF32 data gradient with the mask epilogue + 3xTF32 weight gradient)
    dX, dW, _ = ops.tc_bwd_layer(dZ, X, W, mask=True)
    dX1 = ops.tc_linear(dZ, W, transpose_w=True, mask=X)
    dW1 = ops.tc_wgrad(dZ, X)
    assert _rel(dX, dX1) < 3e-6 and _rel(dW, dW1) < 3e-6
    # deterministic
    dX2, dW2, _ = ops.tc_bwd_layer(dZ, X, W, mask=True)
    assert torch.equal(dX, dX2) and torch.equal(dW, dW2)


@pytest.mark.parametrize("M,D", [(1, 128), (77, 128), (100003, 128), (5000, 64), (3000, 512)])
def test_dot_tail_forward_backward(M, D):
    """Linear(D, 1) as a row dot product (the decoder's last layer, reference models/GNN.py:289-295) and its one-pass
    backward with the ReLU mask of the layer below, against float64."""
    from graphnet_classifier_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + D)
    X = torch.relu(torch.randn(M, D, generator=g, device="cuda"))
    w = torch.randn(1, D, generator=g, device="cuda") / 7
    b = torch.randn(1, generator=g, device="cuda")
    dy = torch.randn(M, 1, generator=g, device="cuda") * 1e-3
    y = ops.dot_tail_fwd(X, w, b)
    assert y.shape == (M, 1) and _rel(y, X.double() @ w.double().t() + b.double()) < 1e-6
    dX, dw, db = ops.dot_tail_bwd(X, w, dy, relu_mask=True)
    ref = (dy.double() @ w.double()) * (X > 0)
    assert _rel(dX, ref) < 1e-6 and bool((dX[X <= 0] == 0).all())
    assert _rel(dw, dy.double().t() @ X.double()) < 2e-6 and _rel(db, dy.double().sum().reshape(1)) < 2e-6
    dX2, _, _ = ops.dot_tail_bwd(X, w, dy, relu_mask=False)
    assert _rel(dX2, dy.double() @ w.double()) < 1e-6


def test_deferred_reduction_equals_immediate():
    """One reduction launch for the layers of a whole pass (ops.DeferredBwdReduce) gives the bits of the per-layer form,
    in overwrite and in accumulate mode, for layers of different row counts (different numbers of partial slices)."""
    from graphnet_classifier_b200 import _lib, ops
    cases = [_case(M, seed=M) for M in (5, 1000, 148 * 32 + 7, 70001)]
    imm = [ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True) for dZ, X, W, _ in cases]
    base = torch.full((128, 384), 0.25, device="cuda")
    acc_i = [(base.clone(), torch.ones(128, device="cuda")) for _ in cases]
    for (dZ, X, W, _), (dWo, dbo) in zip(cases, acc_i):
        ops.tc_bwd_layer(dZ, X, W, dW_out=dWo[:, 128:256], want_db=True, db_out=dbo, accumulate=True)
    _lib.reset_launch_count()
    acc_d = [(base.clone(), torch.ones(128, device="cuda")) for _ in cases]
    with ops.DeferredBwdReduce():
        dfr = [ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True) for dZ, X, W, _ in cases]
        for (dZ, X, W, _), (dWo, dbo) in zip(cases, acc_d):
            ops.tc_bwd_layer(dZ, X, W, dW_out=dWo[:, 128:256], want_db=True, db_out=dbo, accumulate=True)
    torch.cuda.synchronize()
    assert _lib.launch_count() == 2 * len(cases) + 1          # one launch per layer + ONE reduction
    for (dX0, dW0, db0), (dX1, dW1, db1) in zip(imm, dfr):
        assert torch.equal(dX0, dX1) and torch.equal(dW0, dW1) and torch.equal(db0, db1)
    for (w0, b0), (w1, b1) in zip(acc_i, acc_d):
        assert torch.equal(w0, w1) and torch.equal(b0, b1)
        assert bool((w1[:, :128] == 0.25).all()) and bool((w1[:, 256:] == 0.25).all())
