"""The packed CUDA path at the dimensions BASELINE.json's configs 3, 4 and 5 name, against fixtures computed by the
REAL reference modules at those dimensions (oracle/make_golden.py golden_*_cfg*): GAT n ~ 203 / heads [2,2] /
32 -> 64 -> 64 -> 32; DiffPool N = 1000 -> K = 100, D = 96, Linear(164 -> 100); Wave 89 one-hot / D = 64 / 27 clusters.
Forward within 1e-5; every gradient within 1e-5 of the fixture, or -- with the float64 evaluation of the pinned
oracle as the conditioning reference -- as close to float64 as the reference's own fp32 arithmetic (conftest.grad_check)."""
import numpy as np
import pytest
import torch

from conftest import grad_check, rel_err
from golden_util import clone_req, conv_names, convs, dense_adj, eigen_cfg5_operands, gat_layers, load, padded
from oracle import dense_ref as D

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dbl(t):
    return t.double().clone().requires_grad_(True)


def test_gat_config3_dims(cuda):
    from tsg import dense, gat
    d = load("dense_gat_cfg3.npz")
    N, Fi, H, O, L = [int(v) for v in d["dims"]]
    n = int(d["n"])
    model = gat.PackedGatEncoder(Fi, H, O, 2, num_layers=L, num_heads=[2, 2]).to(cuda)
    model.load_state_dict({k[len("param/"):]: v for k, v in d.items() if k.startswith("param/")})
    adj = dense_adj(d["ei"], N)
    csr, _, _ = dense.dense_to_csr(adj.to(cuda), [n], [n], want_eid=True)
    gptr = torch.tensor([0, n], device=cuda)
    readout, out = model(d["x"].to(cuda), csr, gptr, N)
    (readout * d["cot"].to(cuda)).sum().backward()
    assert readout.shape == (1, 32) and model.conv_block[0].attention_0.w.shape == (64, 32)
    assert rel_err(readout, d["readout"]) <= TOL
    assert rel_err(out, d["out"]) <= TOL
    # float64 conditioning reference: the pinned dense oracle in double
    layers = gat_layers(d)
    req = [[dict(w=_dbl(h["w"]), a=_dbl(h["a"])) for h in hs] for _, hs in layers]
    o64 = D.dgat_encoder_readout(padded(d["x"], N).double(), adj.double(), req)
    (o64 * d["cot"].double()).sum().backward()
    g64 = {f"{ln}.attention_{h}.{k}": r[k].grad for (ln, _), rq in zip(layers, req) for h, r in enumerate(rq) for k in ("w", "a")}
    for k, p in model.named_parameters():
        if "grad/" + k in d:
            grad_check(p.grad, d["grad/" + k], g64.get(k), TOL, k)


@pytest.mark.parametrize("tc", [True, False])
def test_diffpool_config4_dims(cuda, tc):
    from tsg import dense, diffpool, ops
    ops.USE_TCGEN05 = tc
    try:
        d = load("dense_diffpool_cfg4.npz")
        N, Fi, H, O, L = [int(v) for v in d["dims"]]
        n = int(d["n"])
        model = diffpool.PackedSoftPoolEncoder(N, Fi, H, O, 2, L, assign_hidden_dim=32, assign_ratio=0.1).to(cuda)
        model.load_state_dict({k[len("param/"):]: v for k, v in d.items() if k.startswith("param/")})
        assert model.assign_dim == 100 and model.assign_pred_modules[0].weight.shape == (100, 164)
        adj = dense_adj(d["ei"], N)
        csr, _, _ = dense.dense_to_csr(adj.to(cuda), [n], [n])
        gptr = torch.tensor([0, n], device=cuda)
        has_pad = torch.tensor([n < N], device=cuda)
        out, aux = model.readout(d["x"].to(cuda), csr, gptr, has_pad, return_aux=True)
        (out * d["cot"].to(cuda)).sum().backward()
        assert rel_err(aux["s"], d["assign"]) <= TOL
        assert rel_err(out, d["readout"]) <= TOL
        assert rel_err(model.map_model(out), d["ypred"]) <= TOL
        # float64 conditioning reference
        p64 = dict(conv=convs(d, "conv_first", "conv_block", "conv_last"),
                   assign_conv=convs(d, "assign_conv_first_modules.0", "assign_conv_block_modules.0", "assign_conv_last_modules.0"),
                   conv_after=convs(d, "conv_first_after_pool.0", "conv_block_after_pool.0", "conv_last_after_pool.0"))
        names = dict(conv=("conv_first", "conv_block", "conv_last"),
                     assign_conv=("assign_conv_first_modules.0", "assign_conv_block_modules.0", "assign_conv_last_modules.0"),
                     conv_after=("conv_first_after_pool.0", "conv_block_after_pool.0", "conv_last_after_pool.0"))
        for k in p64:
            p64[k] = [dict(weight=_dbl(c["weight"]), bias=_dbl(c["bias"])) for c in p64[k]]
        p64["assign_pred.weight"] = _dbl(d["param/assign_pred_modules.0.weight"])
        p64["assign_pred.bias"] = _dbl(d["param/assign_pred_modules.0.bias"])
        o64, _ = D.soft_pool_readout(padded(d["x"], N).double(), adj.double(), [n], p64)
        (o64 * d["cot"].double()).sum().backward()
        fwd_noise = rel_err(d["readout"], o64)          # the reference's fp32 noise on a well-conditioned quantity
        g64 = {"assign_pred_modules.0.weight": p64["assign_pred.weight"].grad, "assign_pred_modules.0.bias": p64["assign_pred.bias"].grad}
        for key, (f, b, l) in names.items():
            for c, nm in zip(p64[key], conv_names(f, b, l, L)):
                g64[nm + ".weight"], g64[nm + ".bias"] = c["weight"].grad, c["bias"].grad
        checked = 0
        for k, p in model.named_parameters():
            if "grad/" + k in d:
                grad_check(p.grad, d["grad/" + k], g64.get(k), TOL, f"tc={tc} {k}", fwd_noise); checked += 1
        assert checked >= 20
    finally:
        ops.USE_TCGEN05 = True


def test_wave_config5_dims(cuda):
    from tsg import dense
    d = load("dense_eigen_cfg5.npz")
    N, Fi, H, O, L = [int(v) for v in d["dims"]]
    n, nc = int(d["n"]), int(d["nc"])
    model = dense.PackedWaveEncoder(Fi, H, O, 2, L, num_pool_matrix=1, num_pool_final_matrix=1,
                                    pool_sizes=[10], pred_hidden_dims=[50]).to(cuda)
    model.load_state_dict({k[len("param/"):]: v for k, v in d.items() if k.startswith("param/")})
    x, adj, ap, P, Pf = eigen_cfg5_operands(d)
    csr_adj, _, _ = dense.dense_to_csr(adj.to(cuda), [n], [n])
    csr_pool, _, _ = dense.dense_to_csr(ap.to(cuda), [nc], [nc])
    P0, _, _ = dense.dense_to_csr(P.to(cuda), [n], [nc], transpose=True)
    Pfc, _, _ = dense.dense_to_csr(Pf.to(cuda), [nc], [1], transpose=True)
    y = model(dense.pack_rows(x.to(cuda), [n]), csr_adj, torch.tensor([0, n], device=cuda), [[P0], [Pfc]], [csr_pool],
              [torch.tensor([0, nc], device=cuda)], torch.tensor([0, 1], device=cuda))
    (y * d["cot"].to(cuda)).sum().backward()
    assert rel_err(y, d["y"]) <= TOL
    p64 = dict(conv=[dict(weight=_dbl(c["weight"]), bias=_dbl(c["bias"])) for c in convs(d, "conv_first", "conv_block", "conv_last")],
               conv_after=[[dict(weight=_dbl(c["weight"]), bias=_dbl(c["bias"]))
                            for c in convs(d, "conv_first_after_pool.0", "conv_block_after_pool.0", "conv_last_after_pool.0")]])
    head = [dict(weight=_dbl(d["param/pred_model.0.weight"]), bias=_dbl(d["param/pred_model.0.bias"])),
            dict(weight=_dbl(d["param/pred_model.2.weight"]), bias=_dbl(d["param/pred_model.2.bias"]))]
    o64 = D.mlp(D.wave_readout(x.double(), adj.double(), [ap.double()], [n], [[nc]], [[P.double()], [Pf.double()]], p64,
                               num_pool_matrix=1, num_pool_final_matrix=1), head)
    (o64 * d["cot"].double()).sum().backward()
    g64 = {"pred_model.0.weight": head[0]["weight"].grad, "pred_model.0.bias": head[0]["bias"].grad,
           "pred_model.2.weight": head[1]["weight"].grad, "pred_model.2.bias": head[1]["bias"].grad}
    for c, nm in zip(p64["conv"], conv_names("conv_first", "conv_block", "conv_last", L)):
        g64[nm + ".weight"], g64[nm + ".bias"] = c["weight"].grad, c["bias"].grad
    for c, nm in zip(p64["conv_after"][0], conv_names("conv_first_after_pool.0", "conv_block_after_pool.0", "conv_last_after_pool.0", L)):
        g64[nm + ".weight"], g64[nm + ".bias"] = c["weight"].grad, c["bias"].grad
    for k, p in model.named_parameters():
        grad_check(p.grad, d["grad/" + k], g64.get(k), TOL, k)
