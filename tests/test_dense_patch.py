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
    assert rel_err(me.weight.grad, wo.grad) <= 2e-5 and rel_err(me.bias.grad, bo.grad) <= 2e-5


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
    assert rel_err(me.w.grad, wo.grad) <= 2e-5 and rel_err(me.a.grad, ao.grad) <= 2e-5


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
