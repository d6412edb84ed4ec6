"""K7 DiffPool parity: per-graph contractions (fp32 SIMT and tcgen05 3xTF32) and the whole
SoftPoolingGcnEncoder forward/backward vs the REAL-reference fixture and the dense oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from golden_util import load
from oracle import dense_ref as D

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _ragged(sizes, kx, ky, seed):
    g = torch.Generator().manual_seed(seed)
    n = int(sum(sizes))
    x = torch.randn(n, kx, generator=g); y = torch.randn(n, ky, generator=g)
    ptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    ref = torch.stack([x[ptr[i]:ptr[i + 1]].double().t() @ y[ptr[i]:ptr[i + 1]].double()
                       for i in range(len(sizes))])
    return x, y, ptr, ref


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("kx,ky", [(100, 196), (5, 29), (128, 256), (64, 64)])
def test_seg_contract(cuda, tc, kx, ky):
    from tsg import ops
    sizes = [269, 0, 1, 31, 32, 33, 1000, 64, 7]
    x, y, ptr, ref = _ragged(sizes, kx, ky, kx + ky)
    c = ops.seg_contract_raw(x.to(cuda), y.to(cuda), ptr.to(cuda), tensor_cores=tc)
    torch.cuda.synchronize()
    assert rel_err(c, ref) <= TOL, f"tensor_cores={tc}"
    assert float(c[1].abs().max()) == 0.0                       # empty graph


def test_seg_contract_softmax_operands_tc(cuda):
    """The operands DiffPool really feeds (softmax rows, sparse-ish activations) through tcgen05."""
    from tsg import ops
    sizes = [300, 150, 269, 40]
    x, y, ptr, _ = _ragged(sizes, 100, 196, 5)
    x = torch.softmax(x * 3, dim=1); y = torch.relu(y)
    ref = torch.stack([x[ptr[i]:ptr[i + 1]].double().t() @ y[ptr[i]:ptr[i + 1]].double() for i in range(4)])
    c = ops.seg_contract_raw(x.to(cuda), y.to(cuda), ptr.to(cuda), tensor_cores=True)
    assert rel_err(c, ref) <= TOL


@pytest.mark.parametrize("kin,m", [(100, 96), (100, 196), (7, 5), (128, 128)])
def test_seg_linear_and_autograd(cuda, kin, m):
    from tsg import ops
    sizes = [100, 0, 3, 64, 129]
    g = torch.Generator().manual_seed(kin)
    n = sum(sizes)
    x = torch.randn(n, kin, generator=g); w = torch.randn(len(sizes), kin, m, generator=g)
    dy = torch.randn(n, m, generator=g)
    ptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    xo, wo = x.clone().double().requires_grad_(True), w.clone().double().requires_grad_(True)
    yo = torch.cat([xo[ptr[i]:ptr[i + 1]] @ wo[i] for i in range(len(sizes))])
    yo.backward(dy.double())
    xg, wg = x.to(cuda).requires_grad_(True), w.to(cuda).requires_grad_(True)
    yg = ops.seg_linear(xg, wg, ptr.to(cuda))
    yg.backward(dy.to(cuda))
    assert rel_err(yg, yo) <= TOL and rel_err(xg.grad, xo.grad) <= TOL and rel_err(wg.grad, wo.grad) <= TOL


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("kin,m", [(196, 100), (100, 196), (32, 256), (36, 4), (256, 64)])
def test_seg_linear_raw_tcgen05_and_simt(cuda, tc, transposed, kin, m):
    """the row-local product on tcgen05 (rows = the MMA's M dimension, 128 at a time; 3xTF32) and on the SIMT kernels, both
    weight layouts, ragged graphs around the 128-row tile incl. empty ones, vs float64"""
    from tsg import _lib, ops
    sizes = [269, 0, 1, 127, 128, 129, 1000, 31, 0, 256]
    g = torch.Generator().manual_seed(kin * 7 + m)
    n = sum(sizes)
    x = torch.randn(n, kin, generator=g)
    w = torch.randn(len(sizes), m, kin, generator=g) if transposed else torch.randn(len(sizes), kin, m, generator=g)
    ptr = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64)
    ref = torch.cat([x[ptr[i]:ptr[i + 1]].double() @ (w[i].double().t() if transposed else w[i].double()) for i in range(len(sizes))])
    prof = {}
    _lib.profile = prof
    try:
        y = ops.seg_linear_raw(x.to(cuda), w.to(cuda), ptr.to(cuda), transposed, tensor_cores=tc)
        torch.cuda.synchronize()
    finally:
        _lib.profile = None
    assert ("tsg_seg_linear_tc" in prof) == tc and ("tsg_seg_linear" in prof) == (not tc)
    assert rel_err(y, ref) <= TOL, f"tensor_cores={tc} transposed={transposed}"


@pytest.mark.parametrize("n,k,m", [(20000, 228, 100), (16385, 64, 32), (33000, 36, 128)])
def test_assignment_linear_softmax_on_tcgen05(cuda, n, k, m):
    """DiffPool's assignment Linear + softmax (encoders.py:366-369) at batch size: tsg_linear_tc (flat tcgen05 product +
    in-place softmax(y + b)) and its input gradient on tcgen05, vs float64; small batches keep the fp32 SIMT kernel"""
    from tsg import _lib, ops
    g = torch.Generator().manual_seed(n + k)
    x = torch.randn(n, k, generator=g); w = torch.randn(k, m, generator=g) / np.sqrt(k); b = torch.randn(m, generator=g)
    dy = torch.randn(n, m, generator=g)
    xo, wo, bo = (t.clone().double().requires_grad_(True) for t in (x, w, b))
    yo = torch.softmax(xo @ wo + bo, dim=1)
    yo.backward(dy.double())
    xg, wg, bg = (t.to(cuda).requires_grad_(True) for t in (x, w, b))
    prof = {}
    _lib.profile = prof
    try:
        yg = ops.linear(xg, wg, bg, ops.LIN_SOFTMAX)
        yg.backward(dy.to(cuda))
    finally:
        _lib.profile = None
    assert len(prof.get("tsg_linear_tc", [])) == 2 and "tsg_linear_fwd" not in prof
    assert rel_err(yg, yo) <= TOL and rel_err(xg.grad, xo.grad) <= TOL
    assert rel_err(wg.grad, wo.grad) <= TOL and rel_err(bg.grad, bo.grad) <= TOL
    # below the size threshold the SIMT epilogue kernel runs, same values within the tolerance
    ys = ops.linear(x[:1000].to(cuda), w.to(cuda), b.to(cuda), ops.LIN_SOFTMAX)
    assert rel_err(ys, yo[:1000]) <= TOL


def _load_diffpool(d, cuda):
    from tsg import diffpool
    N, Fi, H, O, L = [int(v) for v in d["dims"]]
    m = diffpool.PackedSoftPoolEncoder(N, Fi, H, O, 2, L, assign_hidden_dim=8, assign_ratio=0.25).to(cuda)
    m.load_state_dict({k[len("param/"):]: v for k, v in d.items() if k.startswith("param/")})
    return m, N


@pytest.mark.parametrize("tc", [False, True])
def test_diffpool_matches_reference_fixture(cuda, tc):
    from tsg import dense, ops
    ops.USE_TCGEN05 = tc
    try:
        d = load("dense_diffpool.npz")
        model, N = _load_diffpool(d, cuda)
        n = int(d["n"])
        csr, _, _ = dense.dense_to_csr(d["adj"].to(cuda), [n], [n])
        xp = dense.pack_rows(d["x"].to(cuda), [n])
        gptr = torch.tensor([0, n], device=cuda)
        has_pad = torch.tensor([n < N], device=cuda)
        out, aux = model.readout(xp, csr, gptr, has_pad, return_aux=True)
        (out * d["cot"].to(cuda)).sum().backward()
        assert rel_err(aux["s"], d["assign"][0, :n]) <= TOL
        assert rel_err(out, d["readout"]) <= TOL
        ypred = model.map_model(out)
        assert rel_err(ypred, d["ypred"]) <= TOL
        for k, p in model.named_parameters():
            if "grad/" + k in d:
                assert rel_err(p.grad, d["grad/" + k]) <= 1e-5, k
    finally:
        ops.USE_TCGEN05 = True


def test_diffpool_dd_shape_vs_oracle(cuda):
    """DD-shape graphs, the BASELINE config-4 dimensions scaled to N = 300 (K = 30), packed vs the
    oracle run graph by graph on the padded dense wire format."""
    from tsg import dense, diffpool, synth
    B, N, Fi, H, O, L = 4, 300, 16, 32, 32, 3
    corpus = synth.make_corpus("DD", 12, seed=21)
    ids = [g for g in range(12) if corpus.num_nodes(g) <= N][:B]
    adj = torch.zeros(B, N, N); ns = []
    for b, gid in enumerate(ids):
        n = corpus.num_nodes(gid); ns.append(n)
        e0, e1 = corpus.edge_ptr[gid], corpus.edge_ptr[gid + 1]
        adj[b, torch.from_numpy(corpus.col[e0:e1]), torch.from_numpy(corpus.row[e0:e1])] = 1.0
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, N, Fi, generator=g) * 2
    for b, n in enumerate(ns):
        x[b, n:] = 0
    torch.manual_seed(11)
    model = diffpool.PackedSoftPoolEncoder(N, Fi, H, O, 2, L, assign_hidden_dim=H, assign_ratio=0.1)
    def cs(first, block, last):
        return [dict(weight=c.weight.detach().clone().requires_grad_(True), bias=c.bias.detach().clone().requires_grad_(True))
                for c in [first] + list(block) + [last]]
    p = dict(conv=cs(model.conv_first, model.conv_block, model.conv_last),
             assign_conv=cs(model.assign_conv_first_modules[0], model.assign_conv_block_modules[0], model.assign_conv_last_modules[0]),
             conv_after=cs(model.conv_first_after_pool[0], model.conv_block_after_pool[0], model.conv_last_after_pool[0]))
    p["assign_pred.weight"] = model.assign_pred_modules[0].weight.detach().clone().requires_grad_(True)
    p["assign_pred.bias"] = model.assign_pred_modules[0].bias.detach().clone().requires_grad_(True)
    cot = torch.randn(B, 2 * (H * (L - 1) + O), generator=g)
    ref = torch.cat([D.soft_pool_readout(x[b:b + 1], adj[b:b + 1], [ns[b]], p)[0] for b in range(B)])
    (ref * cot).sum().backward()
    model = model.to(cuda)
    csr, _, _ = dense.dense_to_csr(adj.to(cuda), ns, ns)
    gptr = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), device=cuda)
    has_pad = torch.tensor([n < N for n in ns], device=cuda)
    out = model.readout(dense.pack_rows(x.to(cuda), ns), csr, gptr, has_pad)
    (out * cot.to(cuda)).sum().backward()
    assert rel_err(out, ref) <= TOL
    assert rel_err(model.conv_first.weight.grad, p["conv"][0]["weight"].grad) <= 1e-5
    assert rel_err(model.assign_pred_modules[0].weight.grad, p["assign_pred.weight"].grad) <= 1e-5
    assert rel_err(model.conv_first_after_pool[0].weight.grad, p["conv_after"][0]["weight"].grad) <= 1e-5


@pytest.mark.parametrize("k,m", [(164, 100), (128, 64), (96, 300)])
def test_linear_bwd_weight_tensor_core_route(cuda, k, m):
    """GEMM-shaped dW = X^T dY on the tcgen05 contraction (fixed 512-row pseudo-segments, 3xTF32) vs a
    float64 product: fp32-level accuracy, and run-to-run identical."""
    from tsg import ops
    g = torch.Generator().manual_seed(k + m)
    n = 20000 + 123
    x = torch.randn(n, k, generator=g).to(cuda)
    dy = torch.randn(n, m, generator=g).to(cuda)
    dw, db = ops.linear_bwd_weight(x, dy, want_bias=True)
    ref = (x.double().t() @ dy.double())
    assert rel_err(dw, ref) <= 1e-5
    assert rel_err(db, dy.double().sum(0)) <= 1e-5
    dw2, _ = ops.linear_bwd_weight(x, dy, want_bias=False)
    assert torch.equal(dw, dw2)


def test_linkpred_loss_matches_reference_fixture_and_oracle(cuda):
    """K15 vs the fixture produced by the REAL SoftPoolingGcnEncoder.loss(linkpred=True), and vs the oracle on a packed
    batch of DD-shape graphs at K = 100 (sizes that cross the 32-row tile, a 1-node graph, an asymmetric adjacency)."""
    from tsg import dense, ops
    d = load("dense_linkpred.npz")
    n = int(d["n"])
    csr, _, _ = dense.dense_to_csr(d["adj"].to(cuda), [n], [n])
    s = d["assign"][0, :n].to(cuda).clone().requires_grad_(True)
    gptr = torch.tensor([0, n], device=cuda)
    l = ops.linkpred_loss(s, gptr, csr, float(n * n))
    l.backward()
    assert rel_err(l, torch.tensor(float(d["link_loss"]))) <= TOL
    assert rel_err(s.grad, d["dassign"][0, :n]) <= TOL
    # packed batch vs oracle
    rng = np.random.default_rng(5)
    ns = [269, 1, 33, 64, 100, 7]
    B, N, K = len(ns), 300, 100
    adj = np.zeros((B, N, N), np.float32)
    for b, m in enumerate(ns):
        a = (rng.random((m, m)) < 0.03)
        a = np.triu(a, 1); a = (a | a.T).astype(np.float32)
        adj[b, :m, :m] = a
    adj[3, 0, 5] = 1.0; adj[3, 5, 0] = 0.0                     # asymmetric entry
    adj = torch.from_numpy(adj)
    g = torch.Generator().manual_seed(1)
    S = torch.softmax(torch.randn(B, N, K, generator=g) * 2, dim=-1)
    for b, m in enumerate(ns):
        S[b, m:] = 0
    So = S.clone().requires_grad_(True)
    lo = D.link_pred_loss(So, adj, ns)
    (lo * 3.0).backward()
    csr, _, _ = dense.dense_to_csr(adj.to(cuda), ns, ns)
    sp = dense.pack_rows(S.to(cuda), ns).clone().requires_grad_(True)
    gptr = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), device=cuda)
    lg = ops.linkpred_loss(sp, gptr, csr, float(sum(m * m for m in ns)))
    (lg * 3.0).backward()
    assert rel_err(lg, lo) <= TOL
    ref = torch.cat([So.grad[b, :m] for b, m in enumerate(ns)])
    assert rel_err(sp.grad, ref) <= TOL
