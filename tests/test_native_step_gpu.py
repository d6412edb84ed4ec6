"""K14 (tsg_sag_triplet_step_compact: forward + MarginRankingLoss + backward of the whole SAGPool model in one C-ABI
call) against the oracle and against the autograd path it replaces: loss, embeddings and all 18 parameter gradients, with
an injected dropout mask (parity runs inject the mask), and the trainer taking the native path."""
import numpy as np
import pytest
import torch

from conftest import grad_check, rel_err
from oracle import pyg_ref as R
from tsg import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _compact(corpus, dev, coalesced=True):
    from tsg import ops
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    return ops.CompactBatch(t(corpus.node_label, np.int32), t(corpus.row, np.int32), t(corpus.col, np.int32),
                            t(corpus.node_ptr, np.int64), t(corpus.edge_ptr, np.int64), corpus.num_node_labels,
                            int(np.diff(corpus.edge_ptr).max()), coalesced)


def _cotangent_triplets(G):
    """well-conditioned triplets: every hinge active with a large margin, so the loss is smooth around the point"""
    rng = np.random.default_rng(0)
    t = np.stack([rng.integers(0, G, 3 * G), rng.integers(0, G, 3 * G), rng.integers(0, G, 3 * G)], 1)
    return t[(t[:, 0] != t[:, 1]) & (t[:, 0] != t[:, 2])].astype(np.int64)


@pytest.mark.parametrize("shape,G,nhid,C", [("DD", 24, 32, 32), ("PROTEINS", 120, 32, 16), ("DD", 6, 64, 8)])
def test_native_step_matches_oracle_and_autograd(cuda, shape, G, nhid, C):
    from tsg import nn as tnn, ops
    corpus = synth.make_corpus(shape, G, seed=41)
    if shape == "PROTEINS":
        corpus.node_label[:] = np.random.default_rng(1).integers(0, 89, corpus.node_label.shape[0]); corpus.num_node_labels = 89
    b = synth.pack(corpus)
    cb = _compact(corpus, cuda)
    params = R.init_sag_params(corpus.num_node_labels, nhid, C, seed=9)
    model = tnn.PackedSAGNet(corpus.num_node_labels, nhid, C, 0.5, 0.5).to(cuda)
    model.load_state_dict(params)
    model.train()
    trip = torch.from_numpy(_cotangent_triplets(G))
    g = torch.Generator().manual_seed(3)
    mask = (torch.rand(G, nhid, generator=g) >= 0.5).float() * 2.0
    margin = 50.0          # all hinges active: gradient well conditioned (tests/test_sag_gpu.py explains the 1.5 case)
    loss, emb = model.native_step(cb, corpus.node_ptr, trip.to(cuda), margin, dropout_mask=mask.to(cuda), return_emb=True)
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    tnn.check_fused_status()
    # oracle (fp32 and fp64 for conditioning)
    res = {}
    for dt in (torch.float32, torch.float64):
        po = {k: v.clone().to(dt).requires_grad_(True) for k, v in params.items()}
        emb_o = R.sag_net_forward(po, torch.from_numpy(b["x"]).to(dt), torch.from_numpy(b["edge_index"]),
                                  torch.from_numpy(b["batch"]), 0.5, dropout_mask=mask.to(dt))
        loss_o, _, _ = R.triplet_margin_loss(emb_o[trip[:, 0]], emb_o[trip[:, 1]], emb_o[trip[:, 2]], margin)
        loss_o.backward()
        res[dt] = (emb_o.detach(), loss_o.detach(), {k: v.grad for k, v in po.items()})
    assert rel_err(emb, res[torch.float32][0]) <= TOL
    assert rel_err(loss, res[torch.float32][1]) <= TOL
    fwd_noise = max(rel_err(res[torch.float32][0], res[torch.float64][0]), 1e-7)
    for k in grads:
        grad_check(grads[k], res[torch.float32][2][k], res[torch.float64][2][k], TOL, k, fwd_noise)
    # the autograd path (K10 + ops.linear head + K9 through torch.autograd) gives the same numbers
    model.zero_grad(set_to_none=True)
    plan, ptrs = model._level_plan(corpus.node_ptr, cuda)
    zenc = model._encode_compact(cb, plan, ptrs)
    lin = lambda layer, t: ops.linear(t, layer.weight.t().contiguous(), layer.bias)
    h1 = torch.relu(lin(model.lin1, zenc)) * mask.to(cuda)
    emb_a = torch.log_softmax(lin(model.lin3, torch.relu(lin(model.lin2, h1))), dim=-1)
    loss_a, _, _ = ops.triplet_loss(emb_a, trip.to(cuda), margin)
    loss_a.backward()
    assert rel_err(emb, emb_a) <= 2e-6 and rel_err(loss, loss_a) <= 2e-6
    for k, p in model.named_parameters():
        # two fp32 evaluations of the same gradient with different head summation orders (the gate against the ORACLE is
        # above): they agree to a few ulp of the terms, which at 1e-8-sized gradients is ~1e-5 relative
        # (lin3.bias = sum of log-softmax gradients, each of which sums to zero: conditioned like 1e-3, hence the float64 reference)
        grad_check(grads[k], p.grad, res[torch.float64][2][k], 5e-5, "vs autograd " + k, fwd_noise)


def test_trainer_takes_the_native_path_and_learns(cuda):
    from tsg import _lib, nn as tnn
    from tsg.train import TripletTrainer
    corpus = synth.make_corpus("DD", 30, seed=5)
    cb = _compact(corpus, cuda)
    torch.manual_seed(0)
    model = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.0).to(cuda)
    trainer = TripletTrainer(model, lr=5e-3, weight_decay=0.0, margin=1.5)
    trip = torch.from_numpy(synth.sample_triplets(corpus.y, 60, seed=1)).to(cuda)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    calls = _lib.launch_calls
    losses = [float(trainer.step(cb, None, corpus.node_ptr, trip)) for _ in range(25)]
    assert _lib.launch_calls - calls == 25, "one C-ABI call per step"
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert any(not torch.equal(before[k], v) for k, v in model.state_dict().items())
    # dropout on: runs, different masks per step (losses differ for identical inputs and frozen weights)
    model.dropout_ratio = 0.5
    with torch.no_grad():
        pass
    l1 = float(model.native_step(cb, corpus.node_ptr, trip, 1.5)); l2 = float(model.native_step(cb, corpus.node_ptr, trip, 1.5))
    assert l1 != l2


def test_trainer_falls_back_when_a_graph_exceeds_the_executor_limits(cuda, monkeypatch):
    """a graph larger than the per-graph CSR budget: the native step declines, the autograd path takes the batch"""
    from tsg import _lib, nn as tnn, ops
    from tsg.train import TripletTrainer
    corpus = synth.make_corpus("DD", 12, seed=6)
    cb = _compact(corpus, cuda)
    torch.manual_seed(0)
    model = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.0).to(cuda)
    trainer = TripletTrainer(model, lr=1e-3, weight_decay=0.0, margin=1.5)
    trip = torch.from_numpy(synth.sample_triplets(corpus.y, 24, seed=1)).to(cuda)
    assert model.native_step_supported(cb, corpus.node_ptr)
    ref = float(trainer.step(cb, None, corpus.node_ptr, trip))
    monkeypatch.setattr(ops, "GRAPH_CSR_MAX_NODES", int(np.diff(corpus.node_ptr).max()) - 1)
    assert not model.native_step_supported(cb, corpus.node_ptr)
    torch.manual_seed(0)
    model2 = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.0).to(cuda)
    trainer2 = TripletTrainer(model2, lr=1e-3, weight_decay=0.0, margin=1.5)
    got = float(trainer2.step(cb, None, corpus.node_ptr, trip))
    assert abs(got - ref) <= 1e-5 * max(abs(ref), 1.0)


def test_native_halves_equal_the_single_call(cuda):
    """tsg_sag_step_fwd_compact + K9 through autograd + tsg_sag_step_bwd_compact (the all-gather formulation's path)
    == tsg_sag_triplet_step_compact: identical embeddings, loss and gradients (same kernels, same order)."""
    from tsg import nn as tnn, ops
    from tsg.train import TripletTrainer
    corpus = synth.make_corpus("DD", 18, seed=77)
    cb = _compact(corpus, cuda)
    torch.manual_seed(4)
    model = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.0).to(cuda)
    model.train()
    trip = torch.from_numpy(synth.sample_triplets(corpus.y, 40, seed=2)).to(cuda)
    loss1, emb1 = model.native_step(cb, corpus.node_ptr, trip, 1.5, return_emb=True)
    loss1 = loss1.clone()
    g1 = {k: p.grad.clone() for k, p in model.named_parameters()}
    emb2, ctx = model.native_forward(cb, corpus.node_ptr)
    e = emb2.clone().requires_grad_(True)
    loss2, _, _ = ops.triplet_loss(e, trip, 1.5)
    loss2.backward()
    model.native_backward(ctx, e.grad)
    assert torch.equal(emb1, emb2) and torch.equal(loss1, loss2.detach())
    for k, p in model.named_parameters():
        assert torch.equal(g1[k], p.grad), k
    # and the trainer's all-gather step runs on it (world = 1: the gather is the identity)
    tr = TripletTrainer(model, lr=1e-3)
    l = tr.step_allgather(cb, None, corpus.node_ptr, trip)
    assert torch.isfinite(l)
