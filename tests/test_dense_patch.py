"""B2 drop-ins (tsg/dense_patch.py): the patched methods, called with the reference's dense wire
format, against the dense oracle -- including padded rows, isolated nodes and B > 1 node-BN."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import dense_ref as D

REF_DENSE = "/root/reference/Code/sage+gat+diffpool"
TOL = 1e-5


def _wire(B, N, F, seed, isolate=()):
    rng = np.random.default_rng(seed)
    adj = np.zeros((B, N, N), np.float32); ns = []
    for b in range(B):
        n = int(rng.integers(N // 2, N)); ns.append(n)
        for i in range(1, n):
            j = int(rng.integers(0, i)); adj[b, i, j] = adj[b, j, i] = 1
        ex = np.triu(rng.random((n, n)) < 0.2, 1)
        adj[b, :n, :n] = np.maximum(adj[b, :n, :n], (ex | ex.T).astype(np.float32))
    for b, v in isolate:
        adj[b, v, :] = 0; adj[b, :, v] = 0
    x = torch.randn(B, N, F, generator=torch.Generator().manual_seed(seed))
    for b, n in enumerate(ns):
        x[b, n:] = 0
    return x, torch.from_numpy(adj), ns


@pytest.mark.skipif(not os.path.isdir(REF_DENSE), reason="reference checkout not present on this box")
def test_install_patches_the_real_reference_classes():
    torch.Tensor.cuda_backup = torch.Tensor.cuda
    sys.path.insert(0, REF_DENSE)
    try:
        sys.modules.pop("encoders", None)
        import encoders
        from tsg import dense_patch
        before = encoders.GraphConv.forward
        done = dense_patch.install(encoders=encoders)
        assert encoders.GraphConv.forward is dense_patch.graphconv_forward is not before
        assert encoders.GcnEncoderGraph.apply_bn is dense_patch.apply_bn
        assert "GraphConv.forward" in done
        encoders.GraphConv.forward = before
    finally:
        sys.path.remove(REF_DENSE)
        sys.modules.pop("encoders", None)


@pytest.mark.gpu
@pytest.mark.parametrize("add_self", [False, True])
def test_graphconv_and_bn_patch(cuda, add_self):
    from tsg import dense_patch
    x, adj, ns = _wire(3, 20, 7, 1)
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(7, 12, generator=g) * 0.4); b = torch.randn(12, generator=g) * 0.3
    wo, bo = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = D.apply_bn(torch.relu(D.graph_conv(x, adj, wo, bo, add_self=add_self)))
    cot = torch.randn(ref.shape, generator=g)
    (ref * cot).sum().backward()
    me = types.SimpleNamespace(weight=w.to(cuda).requires_grad_(True), bias=b.to(cuda).requires_grad_(True),
                               add_self=add_self, normalize_embedding=True, dropout=0.0)
    out = dense_patch.apply_bn(None, torch.relu(dense_patch.graphconv_forward(me, x.to(cuda), adj.to(cuda))))
    (out * cot.to(cuda)).sum().backward()
    assert rel_err(out, ref) <= TOL
    assert rel_err(me.weight.grad, wo.grad) <= 1e-5 and rel_err(me.bias.grad, bo.grad) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("concat", [True, False])
def test_dgathead_patch_with_padding_and_isolated_nodes(cuda, concat):
    from tsg import dense_patch
    x, adj, ns = _wire(1, 24, 6, 3, isolate=((0, 2),))
    g = torch.Generator().manual_seed(4)
    w = torch.randn(6, 8, generator=g) * 0.5; a = torch.randn(16, 1, generator=g) * 0.5
    x[0, ns[0]:] = torch.randn(24 - ns[0], 6, generator=g) * 0.1        # later layers: padded rows are not zero
    wo, ao = w.clone().requires_grad_(True), a.clone().requires_grad_(True)
    ref = D.dgat_head(x, adj, wo, ao, concat=concat)
    cot = torch.randn(ref.shape, generator=g)
    (ref * cot).sum().backward()
    me = types.SimpleNamespace(w=w.to(cuda).requires_grad_(True), a=a.to(cuda).requires_grad_(True), output_dim=8,
                               leakyRELU_neg_input_slope=0.2, concat=concat, dropout=0.0, training=False)
    out = dense_patch.dgathead_forward(me, x.to(cuda), adj.to(cuda))
    (out * cot.to(cuda)).sum().backward()
    assert out.shape == ref.shape
    assert rel_err(out, ref) <= TOL
    assert rel_err(me.w.grad, wo.grad) <= 1e-5 and rel_err(me.a.grad, ao.grad) <= 1e-5


@pytest.mark.gpu
def test_pool_patch(cuda):
    from tsg import dense_patch
    B, N, Dd = 2, 16, 5
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, N, Dd, generator=g)
    pms = []
    for j in range(2):
        P = torch.zeros(B, N, N)
        for b in range(B):
            for v in range(12):
                P[b, v, v // 4] = float(torch.randn((), generator=g))
        pms.append(P)
    xo = x.clone().requires_grad_(True)
    ref = D.eigen_pool(xo, pms)
    cot = torch.randn(ref.shape, generator=g)
    (ref * cot).sum().backward()
    me = types.SimpleNamespace(num_pool=2, pool_matrices=[p.to(cuda) for p in pms])
    xg = x.to(cuda).requires_grad_(True)
    out = dense_patch.pool_forward(me, xg)
    (out * cot.to(cuda)).sum().backward()
    assert rel_err(out, ref) <= TOL and rel_err(xg.grad, xo.grad) <= TOL


@pytest.mark.gpu
def test_softpool_under_the_patch_trains_the_assignment_tower(cuda):
    """SoftPoolingGcnEncoder feeds the DIFFERENTIABLE pooled adjacency S^T A S into the same
    GraphConv.forward (encoders.py:375,378).  The wire-format drop-in must carry d(adj) back to assign_conv_* /
    assign_pred_*: the forward below is encoders.py:327-406 spelled with the patched methods, checked against the
    fixture produced by the REAL reference module (readout and every parameter gradient)."""
    from golden_util import conv_names, load
    from tsg import dense_patch
    d = load("dense_diffpool.npz")
    N, Fi, H, O, L = [int(v) for v in d["dims"]]
    n = int(d["n"])
    P = {k[len("param/"):]: v.to(cuda).clone().requires_grad_(True) for k, v in d.items() if k.startswith("param/")}

    def conv(name):
        return types.SimpleNamespace(weight=P[name + ".weight"], bias=P[name + ".bias"], add_self=False,
                                     normalize_embedding=True, dropout=0.0)

    def gcn_forward(x, adj, names, mask):
        outs = []
        for i, nm in enumerate(names):
            x = dense_patch.graphconv_forward(conv(nm), x, adj)
            if i < len(names) - 1:
                x = dense_patch.apply_bn(None, torch.relu(x))
            outs.append(x)
        z = torch.cat(outs, dim=2)
        return z * mask if mask is not None else z

    x, adj = d["x"].to(cuda), d["adj"].to(cuda)
    mask = torch.zeros(1, N, 1, device=cuda); mask[0, :n] = 1
    z = gcn_forward(x, adj, conv_names("conv_first", "conv_block", "conv_last", L), mask)
    out0 = z.max(dim=1)[0]
    za = gcn_forward(x, adj, conv_names("assign_conv_first_modules.0", "assign_conv_block_modules.0",
                                        "assign_conv_last_modules.0", L), mask)
    s = torch.softmax(za @ P["assign_pred_modules.0.weight"].t() + P["assign_pred_modules.0.bias"], dim=-1) * mask
    xp = s.transpose(1, 2) @ z
    ap = s.transpose(1, 2) @ adj @ s
    assert ap.requires_grad
    z2 = gcn_forward(xp, ap, conv_names("conv_first_after_pool.0", "conv_block_after_pool.0", "conv_last_after_pool.0", L), None)
    out = torch.cat([out0, z2.max(dim=1)[0]], dim=1)
    (out * d["cot"].to(cuda)).sum().backward()
    assert rel_err(out, d["readout"]) <= TOL
    checked = 0
    for k, p in P.items():
        if "grad/" + k in d:
            assert p.grad is not None, f"{k}: no gradient reached it through the patched GraphConv"
            assert rel_err(p.grad, d["grad/" + k]) <= 1e-5, k
            checked += k.startswith("assign_")
    assert checked >= 2 * L + 2
    assert len(dense_patch.CACHE.items) <= dense_patch.CACHE.slots
    assert not any(m.requires_grad for m, *_ in dense_patch.CACHE.items), "a differentiable adjacency was cached"
