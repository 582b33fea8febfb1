"""GPU: the chained-MLP tensor-core kernel (csrc/tc_chain.cu: CTA pairs, bf16x3 operands, hidden
activations kept in tensor memory) against float64, per tail and addend form, at the north_star
tolerance (1e-5)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _maxrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _mlp64(A, layers, pre=None):
    z = A.double() @ layers[0][0].double().t() + layers[0][1].double()
    if pre is not None:
        z = z + pre
    for W, b in layers[1:]:
        z = torch.relu(z) @ W.double().t() + b.double()
    return z


def _layers(gen, n, scale=11.0):
    return [(torch.randn(128, 128, generator=gen) / scale, torch.randn(128, generator=gen) * 0.3) for _ in range(n)]


@pytest.mark.parametrize("M", [1, 127, 256, 257, 1000, 256 * 74 + 131, 256 * 74 * 3 + 5])
@pytest.mark.parametrize("nlayers", [2, 3])
def test_chain_layernorm_residual(libgnc, M, nlayers):
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M * 7 + nlayers)
    A = torch.randn(M, 128, generator=gen) * 2
    layers = _layers(gen, nlayers)
    gamma, beta = torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.2
    res = torch.randn(M, 128, generator=gen)
    cl = [(W.cuda(), b.cuda()) for W, b in layers]
    z = _mlp64(A, layers)
    ref = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5) + res.double()
    got = ops.tc_mlp_chain(A.cuda(), cl, gamma=gamma.cuda(), beta=beta.cuda(), eps=1e-5, residual=res.cuda())
    assert _rel(got, ref) < 3e-6 and _maxrel(got, ref) < RTOL
    # no LayerNorm, no residual: the raw chain
    got = ops.tc_mlp_chain(A.cuda(), cl)
    assert _rel(got, z) < 3e-6 and _maxrel(got, z) < RTOL


def test_chain_gathered_addends_weight_slices_and_table_residual(libgnc):
    """The edge-processor form: z0 = e Wc^T + b + P[src] + Q[dst] with Wc a column slice (ldw = 384);
    residual through a table lookup (the block-0 edge-class form)."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(5)
    M, R = 5000, 900
    A = torch.randn(M, 128, generator=gen)
    W0 = torch.randn(128, 384, generator=gen) / 20
    b0 = torch.randn(128, generator=gen) * 0.1
    rest = _layers(gen, 2)
    P, Q = torch.randn(R, 128, generator=gen), torch.randn(R, 128, generator=gen)
    i0, i1 = torch.randint(0, R, (M,), generator=gen), torch.randint(0, R, (M,), generator=gen)
    gamma, beta = torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.2
    tab, ti = torch.randn(4, 128, generator=gen), torch.randint(0, 4, (M,), generator=gen)
    W0c = W0.cuda()
    layers = [(W0[:, 256:384], b0)] + rest
    cl = [(W0c[:, 256:384], b0.cuda())] + [(W.cuda(), b.cuda()) for W, b in rest]
    z = _mlp64(A, layers, pre=P.double()[i0] + Q.double()[i1])
    ln = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5)
    got = ops.tc_mlp_chain(A.cuda(), cl, gather0=(P.cuda(), i0.int().cuda()), gather1=(Q.cuda(), i1.int().cuda()),
                           gamma=gamma.cuda(), beta=beta.cuda(), residual=A.cuda())
    assert _maxrel(got, ln + A.double()) < RTOL
    got = ops.tc_mlp_chain(A.cuda(), cl, gather0=(P.cuda(), i0.int().cuda()), gather1=(Q.cuda(), i1.int().cuda()),
                           gamma=gamma.cuda(), beta=beta.cuda(), residual=(tab.cuda(), ti.int().cuda()))
    assert _maxrel(got, ln + tab.double()[ti]) < RTOL
    # the node-processor form: one plain addend (identity index), residual = another tensor
    T, h = torch.randn(M, 128, generator=gen), torch.randn(M, 128, generator=gen)
    z = _mlp64(A, layers, pre=T.double())
    ln = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5)
    got = ops.tc_mlp_chain(A.cuda(), cl, gather0=(T.cuda(), None), gamma=gamma.cuda(), beta=beta.cuda(), residual=h.cuda())
    assert _maxrel(got, ln + h.double()) < RTOL
    got = ops.tc_mlp_chain(A.cuda(), cl, gather1=(T.cuda(), None), gamma=gamma.cuda(), beta=beta.cuda())
    assert _maxrel(got, ln) < RTOL


@pytest.mark.parametrize("M", [77, 256 * 80 + 3])
def test_chain_decoder_dot_tail(libgnc, M):
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M)
    A = torch.randn(M, 128, generator=gen)
    layers = _layers(gen, 2)
    w, b = torch.randn(1, 128, generator=gen) / 11, torch.randn(1, generator=gen)
    cl = [(W.cuda(), bb.cuda()) for W, bb in layers]
    ref = torch.relu(_mlp64(A, layers)) @ w.double().t() + b.double()
    got = ops.tc_mlp_chain(A.cuda(), cl, dot_w=w.cuda(), dot_b=b.cuda())
    assert got.shape == (M, 1) and _maxrel(got, ref) < RTOL


def test_chain_matches_per_layer_engine_at_scale(libgnc):
    """2 M rows (every CTA pair runs many tiles): chained result == the per-layer 3xTF32 engine to 1e-5."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(11)
    M = 256 * 74 * 110 + 19
    A = torch.randn(M, 128, generator=gen).cuda()
    layers = [(W.cuda(), b.cuda()) for W, b in _layers(gen, 3)]
    gamma, beta = (torch.rand(128, generator=gen) + 0.5).cuda(), (torch.randn(128, generator=gen) * 0.2).cuda()
    a1 = ops.tc_linear(A, layers[0][0], bias=layers[0][1], relu=True)
    a2 = ops.tc_linear(a1, layers[1][0], bias=layers[1][1], relu=True)
    ref = ops.tc_linear(a2, layers[2][0], bias=layers[2][1], gamma=gamma, beta=beta, eps=1e-5, residual=A)
    got = ops.tc_mlp_chain(A, layers, gamma=gamma, beta=beta, eps=1e-5, residual=A)
    assert _maxrel(got, ref) < RTOL


@pytest.mark.parametrize("M", [300, 256 * 74 * 2 + 77, 256 * 74 * 5 + 256 * 30 + 9])
def test_chain_prestage_table_form(libgnc, M):
    """Block 0 of the grid-graph path: first operand relu(R[class] + P[src] + Q[dst] + b0) built in the launch,
    two layers, LayerNorm, residual through the class table."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M)
    Rn = 500
    R, e_tab = torch.randn(4, 128, generator=gen), torch.randn(4, 128, generator=gen)
    P, Q = torch.randn(Rn, 128, generator=gen), torch.randn(Rn, 128, generator=gen)
    cls = torch.randint(0, 4, (M,), generator=gen)
    i0, i1 = torch.randint(0, Rn, (M,), generator=gen), torch.randint(0, Rn, (M,), generator=gen)
    b0 = torch.randn(128, generator=gen) * 0.2
    layers = _layers(gen, 2)
    gamma, beta = torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.2
    a1 = torch.relu(R.double()[cls] + P.double()[i0] + Q.double()[i1] + b0.double())
    z = torch.relu(a1 @ layers[0][0].double().t() + layers[0][1].double()) @ layers[1][0].double().t() + layers[1][1].double()
    ref = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5) + e_tab.double()[cls]
    got = ops.tc_mlp_chain(None, [(W.cuda(), b.cuda()) for W, b in layers], pre=(R.cuda(), cls.int().cuda(), b0.cuda()),
                           gather0=(P.cuda(), i0.int().cuda()), gather1=(Q.cuda(), i1.int().cuda()),
                           gamma=gamma.cuda(), beta=beta.cuda(), residual=(e_tab.cuda(), cls.int().cuda()))
    assert _maxrel(got, ref) < RTOL


@pytest.mark.parametrize("tiles_per_pair", [1, 2, 3, 6, 7])
@pytest.mark.parametrize("form", ["edge", "node"])
def test_two_tile_kernel_many_tiles(libgnc, tiles_per_pair, form):
    """The two-tiles-in-flight kernel (edge / node processor shapes) with odd and even tile counts per CTA
    pair and a ragged last tile, against the per-layer 3xTF32 engine and the single-tile kernel."""
    import os
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(tiles_per_pair * 3 + len(form))
    M = 256 * 74 * (tiles_per_pair - 1) + 256 * 40 + 57        # some pairs run one tile more than others
    R = 3000
    A = torch.randn(M, 128, generator=gen).cuda()
    layers = [(W.cuda(), b.cuda()) for W, b in _layers(gen, 3)]
    gamma, beta = (torch.rand(128, generator=gen) + 0.5).cuda(), (torch.randn(128, generator=gen) * 0.2).cuda()
    if form == "edge":
        P, Q = torch.randn(R, 128, generator=gen).cuda(), torch.randn(R, 128, generator=gen).cuda()
        i0 = torch.randint(0, R, (M,), generator=gen).int().cuda()
        i1 = torch.randint(0, R, (M,), generator=gen).int().cuda()
        kw = dict(gather0=(P, i0), gather1=(Q, i1))
        a1 = ops.tc_linear(A, layers[0][0], bias=layers[0][1], gather0=(P, i0), gather1=(Q, i1), relu=True)
    else:
        T = torch.randn(M, 128, generator=gen).cuda()
        kw = dict(gather0=(T, None))
        a1 = ops.tc_linear(A, layers[0][0], bias=layers[0][1], addend=T, relu=True)
    a2 = ops.tc_linear(a1, layers[1][0], bias=layers[1][1], relu=True)
    ref = ops.tc_linear(a2, layers[2][0], bias=layers[2][1], gamma=gamma, beta=beta, eps=1e-5, residual=A)
    got = ops.tc_mlp_chain(A, layers, gamma=gamma, beta=beta, eps=1e-5, residual=A, **kw)
    assert _maxrel(got, ref) < RTOL
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("tiles_per_pair", [1, 2, 5])
@pytest.mark.parametrize("form", ["edge", "node"])
def test_two_tile_kernel_training_stash(libgnc, tiles_per_pair, form):
    """Training forward of a block MLP: the chained launch also writes a1, a2, the LayerNorm input and the row statistics
    (what the backward reads) - against float64 and against the per-layer engine's saved tensors; the output must be
    the bits of the same launch without the stash.  Row pitch of the stash wider than 128; ragged last tile."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(tiles_per_pair * 5 + len(form))
    M = 256 * 74 * (tiles_per_pair - 1) + 256 * 33 + 101
    R = 2500
    A = torch.randn(M, 128, generator=gen).cuda()
    layers = [(W.cuda(), b.cuda()) for W, b in _layers(gen, 3)]
    gamma, beta = (torch.rand(128, generator=gen) + 0.5).cuda(), (torch.randn(128, generator=gen) * 0.2).cuda()
    if form == "edge":
        P, Q = torch.randn(R, 128, generator=gen).cuda(), torch.randn(R, 128, generator=gen).cuda()
        i0 = torch.randint(0, R, (M,), generator=gen).int().cuda()
        i1 = torch.randint(0, R, (M,), generator=gen).int().cuda()
        kw = dict(gather0=(P, i0), gather1=(Q, i1))
        pre = P[i0.long()].double() + Q[i1.long()].double()
    else:
        T = torch.randn(M, 128, generator=gen).cuda()
        kw = dict(gather0=(T, None))
        pre = T.double()
    back = torch.full((3, M, 160), float("nan"), device="cuda")          # pitch 160: the stash honours ld_stash
    a1, a2, z = back[0, :, :128], back[1, :, :128], back[2, :, :128]
    mean, rstd = torch.full((M,), float("nan"), device="cuda"), torch.full((M,), float("nan"), device="cuda")
    got = ops.tc_mlp_chain(A, layers, gamma=gamma, beta=beta, eps=1e-5, residual=A, stash=(a1, a2, z, mean, rstd), **kw)
    plain = ops.tc_mlp_chain(A, layers, gamma=gamma, beta=beta, eps=1e-5, residual=A, **kw)
    assert torch.equal(got, plain)
    assert torch.isnan(back[:, :, 128:]).all()                           # nothing written beside the rows
    z0 = A.double() @ layers[0][0].double().t() + layers[0][1].double() + pre
    r1 = torch.relu(z0)
    r2 = torch.relu(r1 @ layers[1][0].double().t() + layers[1][1].double())
    z3 = r2 @ layers[2][0].double().t() + layers[2][1].double()
    assert _maxrel(a1, r1) < RTOL and _maxrel(a2, r2) < RTOL and _maxrel(z, z3) < RTOL
    mu = z3.mean(1)
    rs = 1.0 / torch.sqrt(z3.var(1, unbiased=False) + 1e-5)
    assert _maxrel(mean, mu) < RTOL and _maxrel(rstd, rs) < RTOL
    # the ReLU masks the backward derives from a1 / a2 are those of the per-layer engine wherever the float64
    # pre-activation is not within rounding of zero
    safe = (z0.abs() > 1e-4).cuda()
    assert bool((((a1 > 0) == (z0 > 0).cuda()) | ~safe).all())
    # and the LayerNorm output follows from the stashed tensors
    y = (z.double() - mean.double()[:, None]) * rstd.double()[:, None] * gamma.double() + beta.double() + A.double()
    assert _maxrel(got, y) < RTOL
    with pytest.raises(Exception):                                       # the stash belongs to the 3-layer block forms
        ops.tc_mlp_chain(A, layers[:2], gamma=gamma, beta=beta, residual=A, stash=(a1, a2, z, mean, rstd))


def test_chain_out_of_domain_is_loud(libgnc):
    """fp16 two-piece operands cover |activation| < 4094 (include/gnc.h): beyond it the result must be
    non-finite, never a finite wrong number."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(1)
    A = torch.randn(300, 128, generator=gen)
    A[7, 5] = 6000.0
    layers = [(W.cuda(), b.cuda()) for W, b in _layers(gen, 2)]
    out = ops.tc_mlp_chain(A.cuda(), layers)
    assert not torch.isfinite(out[7]).all()
    ok = torch.ones(300, dtype=torch.bool); ok[7] = False
    assert torch.isfinite(out[ok.cuda()]).all()


@pytest.mark.parametrize("M", [200, 256 * 74 * 4 + 256 * 11 + 3])
def test_chain_two_operand_node_form(libgnc, M):
    """NodeProcessor form without a T tensor: z0 = agg Vb^T + h Va^T + c0 (cat([x, agg]) @ V0^T, models/GNN.py:100),
    two more layers, LayerNorm, + h."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M + 1)
    agg, h = torch.randn(M, 128, generator=gen) * 2, torch.randn(M, 128, generator=gen)
    V0 = torch.randn(128, 256, generator=gen) / 16
    c0 = torch.randn(128, generator=gen) * 0.2
    rest = _layers(gen, 2)
    gamma, beta = torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.2
    z = torch.relu(torch.cat([h, agg], 1).double() @ V0.double().t() + c0.double())
    z = torch.relu(z @ rest[0][0].double().t() + rest[0][1].double()) @ rest[1][0].double().t() + rest[1][1].double()
    ref = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5) + h.double()
    V0c = V0.cuda()
    got = ops.tc_mlp_chain(agg.cuda(), [(V0c[:, 128:256], c0.cuda())] + [(W.cuda(), b.cuda()) for W, b in rest],
                           operand2=(h.cuda(), V0c[:, 0:128]), gamma=gamma.cuda(), beta=beta.cuda(), residual=h.cuda())
    assert _maxrel(got, ref) < RTOL


@pytest.mark.parametrize("M", [1, 200, 256 * 74 + 77, 256 * 74 * 4 + 256 * 11 + 3, 256 * 74 * 5 + 31])
def test_chain_node_form_with_aggregating_loader(libgnc, M):
    """NodeProcessor form with scatter_sum folded into the launch: the loader forms agg[m] = e[eid0] + e[eid1] from the
    edge table (in-degree 0, 1 or 2 per node, edge ids in any order) - the bits of the aggregation kernel followed by the
    two-operand launch, odd / even tile counts per CTA pair, ragged last tile."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M + 5)
    deg = torch.randint(0, 3, (M,), generator=gen)
    if M > 2:
        deg[0], deg[M - 1] = 2, 1
    E = int(deg.sum())
    rowptr = torch.zeros(M + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(deg, 0).int()
    eid = torch.randperm(max(E, 1), generator=gen)[:E].int()        # a permutation: every edge row has one destination
    e = (torch.randn(max(E, 1), 128, generator=gen) * 2).cuda()
    h = torch.randn(M, 128, generator=gen).cuda()
    V0 = (torch.randn(128, 256, generator=gen) / 16).cuda()
    c0 = (torch.randn(128, generator=gen) * 0.2).cuda()
    rest = [(W.cuda(), b.cuda()) for W, b in _layers(gen, 2)]
    gamma, beta = (torch.rand(128, generator=gen) + 0.5).cuda(), (torch.randn(128, generator=gen) * 0.2).cuda()
    rp, ei = rowptr.cuda(), eid.cuda()
    agg = ops._agg_raw(rp, ei, e, M)
    layers = [(V0[:, 128:256], c0)] + rest
    ref = ops.tc_mlp_chain(agg, layers, operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta, residual=h)
    got = ops.tc_mlp_chain(e, layers, operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta, residual=h, agg=(rp, ei))
    assert got.shape == (M, 128) and torch.isfinite(got).all()
    assert torch.equal(got, ref)
    # and against float64 from the edge table itself
    a64 = torch.zeros(M, 128, dtype=torch.float64)
    dst = torch.repeat_interleave(torch.arange(M), deg)
    a64.index_add_(0, dst, e.cpu().double()[eid.long()])
    z = torch.relu(torch.cat([h.cpu().double(), a64], 1) @ V0.cpu().double().t() + c0.cpu().double())
    z = torch.relu(z @ rest[0][0].cpu().double().t() + rest[0][1].cpu().double()) @ rest[1][0].cpu().double().t() + rest[1][1].cpu().double()
    want = torch.nn.functional.layer_norm(z, (128,), gamma.cpu().double(), beta.cpu().double(), 1e-5) + h.cpu().double()
    assert _maxrel(got, want) < RTOL


@pytest.mark.parametrize("M,k", [(500, 3), (256 * 74 * 3 + 100, 3), (1000, 8), (700, 1)])
def test_chain_narrow_first_layer(libgnc, M, k):
    """Encoder form: relu(x Wn^T + bn) built by the loader from k <= 8 input columns, two layers, LayerNorm."""
    from graphnet_classifier_b200 import ops
    gen = torch.Generator().manual_seed(M + k)
    x = torch.rand(M, k, generator=gen) * 255
    Wn, bn = (torch.rand(128, k, generator=gen) - 0.5) / k ** 0.5, torch.randn(128, generator=gen) * 0.3
    rest = _layers(gen, 2)
    gamma, beta = torch.rand(128, generator=gen) + 0.5, torch.randn(128, generator=gen) * 0.2
    a1 = torch.relu(x.double() @ Wn.double().t() + bn.double())
    z = torch.relu(a1 @ rest[0][0].double().t() + rest[0][1].double()) @ rest[1][0].double().t() + rest[1][1].double()
    ref = torch.nn.functional.layer_norm(z, (128,), gamma.double(), beta.double(), 1e-5)
    got = ops.tc_mlp_chain(x.cuda(), [(W.cuda(), b.cuda()) for W, b in rest], narrow=(Wn.cuda(), bn.cuda()),
                           gamma=gamma.cuda(), beta=beta.cuda())
    assert _maxrel(got, ref) < RTOL
