"""K4 dense-GAT parity: CUDA path vs the fixture produced by the REAL reference DGATEncoderGraph and
vs the dense oracle on a packed batch that includes isolated nodes, a full (n == N) graph and ties."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from golden_util import gat_layers, load
from oracle import dense_ref as D

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _load_gat(d, cuda, heads):
    from tsg import gat
    N, Fi, H, O, L = [int(v) for v in d["dims"]]
    m = gat.PackedGatEncoder(Fi, H, O, 2, num_layers=L, num_heads=heads).to(cuda)
    m.load_state_dict({k[len("param/"):]: v for k, v in d.items() if k.startswith("param/")})
    return m, N


def test_gat_matches_reference_fixture(cuda):
    from tsg import dense
    d = load("dense_gat.npz")
    model, N = _load_gat(d, cuda, [2, 2, 2])
    n = int(d["n"])
    csr, _, _ = dense.dense_to_csr(d["adj"].to(cuda), [n], [n], want_eid=True)
    xp = dense.pack_rows(d["x"].to(cuda), [n])
    gptr = torch.tensor([0, n], device=cuda)
    readout, out = model(xp, csr, gptr, N)
    (readout * d["cot"].to(cuda)).sum().backward()
    assert rel_err(readout, d["readout"]) <= TOL
    assert rel_err(out, d["out"]) <= TOL
    for k, p in model.named_parameters():
        if "grad/" + k in d:
            assert rel_err(p.grad, d["grad/" + k]) <= 1e-5, k


def test_gat_packed_batch_with_isolated_nodes_vs_oracle(cuda):
    from tsg import dense, gat
    rng = np.random.default_rng(3)
    B, N, Fi, H, O, L = 10, 30, 6, 8, 8, 3
    adj = np.zeros((B, N, N), np.float32); ns = []
    for b in range(B):
        n = int(rng.integers(3, N + 1)) if b != 2 else N
        ns.append(n)
        for i in range(1, n):
            j = int(rng.integers(0, i)); adj[b, i, j] = adj[b, j, i] = 1
        ex = np.triu(rng.random((n, n)) < 0.2, 1)
        adj[b, :n, :n] = np.maximum(adj[b, :n, :n], (ex | ex.T).astype(np.float32))
    # isolate two real nodes of graph 1 and one of graph 4 (rows + columns)
    for b, v in ((1, 0), (1, ns[1] - 1), (4, 1)):
        adj[b, v, :] = 0; adj[b, :, v] = 0
    adj = torch.from_numpy(adj)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, N, Fi, generator=g)
    for b, n in enumerate(ns):
        x[b, n:] = 0
    torch.manual_seed(9)
    model = gat.PackedGatEncoder(Fi, H, O, 2, num_layers=L, num_heads=[3, 3, 3])
    layers_o = [[dict(w=h.w.detach().clone().requires_grad_(True), a=h.a.detach().clone().requires_grad_(True))
                 for h in layer.heads()] for layer in model.layers()]
    cot = torch.randn(B, O, generator=g)
    ref = torch.cat([D.dgat_encoder_readout(x[b:b + 1], adj[b:b + 1], layers_o) for b in range(B)])
    (ref * cot).sum().backward()
    model = model.to(cuda)
    csr, _, _ = dense.dense_to_csr(adj.to(cuda), ns, ns, want_eid=True)
    gptr = torch.tensor(np.concatenate([[0], np.cumsum(ns)]), device=cuda)
    out = model.readout(dense.pack_rows(x.to(cuda), ns), csr, gptr, N)
    (out * cot.to(cuda)).sum().backward()
    assert rel_err(out, ref) <= TOL
    for layer, lo in zip(model.layers(), layers_o):
        for h, o in zip(layer.heads(), lo):
            assert rel_err(h.w.grad, o["w"].grad) <= 1e-5
            assert rel_err(h.a.grad, o["a"].grad) <= 1e-5
