"""K10 native step executor == the op-by-op PackedSAGNet path (same kernels, same order): outputs and
every parameter gradient bit-identical; and both within 1e-5 of the CPU oracle through test_sag_gpu."""
import numpy as np
import pytest
import torch

from tsg import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,G,nhid", [("DD", 12, 32), ("PROTEINS", 50, 32), ("DD", 6, 128), ("JANY", 5, 16)])
def test_executor_matches_op_by_op(cuda, shape, G, nhid):
    from tsg import nn as tnn
    c = synth.make_corpus(shape, G, seed=5)
    b = synth.pack(c)
    x = torch.from_numpy(b["x"]).to(cuda)
    ei = torch.from_numpy(b["edge_index"]).to(cuda)
    torch.manual_seed(3)
    model = tnn.PackedSAGNet(c.num_node_labels, nhid, 8, 0.5, 0.0).to(cuda)      # dropout 0: deterministic head
    cot = torch.randn(G, 8, generator=torch.Generator().manual_seed(1)).to(cuda)
    res = {}
    for use in (False, True):
        tnn.USE_EXECUTOR = use
        model.zero_grad(set_to_none=True)
        out = model(x, ei, b["node_ptr"])
        (out * cot).sum().backward()
        res[use] = (out.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    tnn.USE_EXECUTOR = True
    assert torch.equal(res[True][0], res[False][0])
    for k in res[False][1]:
        if (k.endswith("score_layer.weight") or k.endswith("score_layer.bias")) and nhid % 4 == 0:
            # the executor's fused level backward sums h^T dsw (k_sag_conv_bwd_v4) and sum(dscore)
            # (k_gate_score_bwd) in its own fixed order
            scale = float(res[False][1][k].abs().max()) + 1e-12
            assert float((res[True][1][k] - res[False][1][k]).abs().max()) <= 2e-5 * scale, k
        else:
            assert torch.equal(res[True][1][k], res[False][1][k]), k


def test_executor_single_node_graphs(cuda):
    """graphs of one node (squeeze() edge case of layers.py:18) and edge-free graphs."""
    from tsg import nn as tnn
    G = 9
    node_ptr = np.arange(G + 1, dtype=np.int64)
    x = torch.eye(4)[torch.arange(G) % 4].to(cuda)
    ei = torch.zeros(2, 0, dtype=torch.int64, device=cuda)
    torch.manual_seed(0)
    model = tnn.PackedSAGNet(4, 8, 3, 0.5, 0.0).to(cuda)
    outs = []
    for use in (False, True):
        tnn.USE_EXECUTOR = use
        outs.append(model(x, ei, node_ptr).detach().clone())
    tnn.USE_EXECUTOR = True
    assert torch.equal(outs[0], outs[1])


def test_host_feeders_agree(cuda):
    """fp32-wire feeder (x / edge_index from the host), compact feeder with K0 expansion and the blocking per-step call
    produce identical losses: the inputs that reach the step are the same tensors.  The compact feeder's direct path
    (labels + local endpoints into the executor, nothing expanded) has the identical first loss -- its forward is
    bit-identical -- and stays within fp32 summation-order distance afterwards (conv1's dW is a segment sum there)."""
    import copy
    from tsg import nn as tnn
    from tsg.train import TripletTrainer
    corpus = synth.make_corpus("DD", 20, seed=3)
    batches, compact = [], []
    for b in range(3):
        trip = synth.sample_triplets(corpus.y, 20, seed=b)
        ids = np.concatenate([trip[:, 0], trip[:, 1], trip[:, 2]])
        pk = synth.pack(corpus, ids)
        tidx = torch.from_numpy(np.stack([np.arange(20), 20 + np.arange(20), 40 + np.arange(20)], 1).astype(np.int64))
        batches.append(dict(x=torch.from_numpy(pk["x"]).pin_memory(), edge_index=torch.from_numpy(pk["edge_index"]).pin_memory(),
                            node_ptr=pk["node_ptr"], triplets=tidx.pin_memory()))
        sel = synth.select(corpus, ids)
        compact.append(dict(label=torch.from_numpy(sel.node_label.astype(np.int32)).pin_memory(),
                            row=torch.from_numpy(sel.row.astype(np.int32)).pin_memory(),
                            col=torch.from_numpy(sel.col.astype(np.int32)).pin_memory(),
                            node_ptr=sel.node_ptr.copy(), edge_ptr=sel.edge_ptr.copy(), triplets=tidx.pin_memory()))
    torch.manual_seed(1)
    base = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.0).to(cuda)
    res = []
    for mode in ("wire", "compact", "blocking", "direct"):
        tr = TripletTrainer(copy.deepcopy(base))
        if mode == "wire":
            res.append(tr.run_from_host(batches, cuda))
        elif mode == "compact":
            res.append(tr.run_from_host_compact(compact, cuda, corpus.num_node_labels, expand=True))
        elif mode == "direct":
            res.append(tr.run_from_host_compact(compact, cuda, corpus.num_node_labels))
        else:
            res.append([tr.step_from_host(b["x"], b["edge_index"], b["node_ptr"], b["triplets"], cuda) for b in batches])
    assert res[0] == res[1] == res[2] and len(res[0]) == 3
    assert res[3][0] == res[0][0]
    np.testing.assert_allclose(res[3], res[0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N,F", [(5000, 32), (333, 128), (70000, 16), (1, 4)])
def test_fused_level_backward_matches_the_kernel_sequence(cuda, N, F):
    """tsg_sag_conv_bwd_fused == tsg_gate_gather_bwd -> tsg_relu_bwd_colsum_rank1 (dhm, dbias bit-identical) and
    tsg_linear_bwd_weight(h, dsw) (dws to fp32 summation order; fp64 reference), deterministic run to run."""
    from tsg._lib import call, ptr, stream_ptr, lib
    g = torch.Generator().manual_seed(N + F)
    k = max(N // 2, 1)
    perm = torch.randperm(N, generator=g)[:k]
    inv = torch.full((N,), -1, dtype=torch.int32); inv[perm] = torch.arange(k, dtype=torch.int32)
    dxo = torch.randn(k, F, generator=g).to(cuda); h = torch.relu(torch.randn(N, F, generator=g)).to(cuda)
    score = torch.randn(N, generator=g).to(cuda); dsw = torch.randn(N, generator=g).to(cuda)
    wsv = torch.randn(F, generator=g).to(cuda); inv = inv.to(cuda)
    wsb = max(2 * lib.tsg_colsum_workspace_bytes(N, F), lib.tsg_linear_bwd_weight_workspace_bytes(F, 1))
    ws = torch.empty(wsb, dtype=torch.uint8, device=cuda)
    dh = torch.empty(N, F, device=cuda); dscore = torch.empty(N, device=cuda)
    call("tsg_gate_gather_bwd", ptr(dxo), ptr(h), ptr(score), ptr(inv), ptr(dh), ptr(dscore), N, F, stream_ptr())
    dscore2 = torch.empty(N, device=cuda)
    call("tsg_gate_gather_bwd", ptr(dxo), ptr(h), ptr(score), ptr(inv), None, ptr(dscore2), N, F, stream_ptr())
    assert torch.equal(dscore, dscore2)                                  # dscore-only mode
    # perm-driven score half: same dscore, plus its sum
    wsg = torch.empty(lib.tsg_gate_score_bwd_workspace_bytes(), dtype=torch.uint8, device=cuda)
    perm_d = perm.to(cuda)
    for _ in range(2):
        dscore3 = torch.full((N,), 7.0, device=cuda); dbs = torch.empty(1, device=cuda)
        call("tsg_gate_score_bwd", ptr(dxo), ptr(h), ptr(score), ptr(perm_d), k, N, F, ptr(dscore3), ptr(dbs), ptr(wsg),
             wsg.numel(), stream_ptr())
        assert torch.equal(dscore, dscore3)
        np.testing.assert_allclose(float(dbs), float(dscore.double().sum()), rtol=1e-5, atol=1e-4)
    dhm = torch.empty(N, F, device=cuda); db = torch.empty(F, device=cuda)
    call("tsg_relu_bwd_colsum_rank1", ptr(dh), ptr(h), ptr(dsw), ptr(wsv), ptr(dhm), ptr(db), N, F, ptr(ws), wsb, stream_ptr())
    outs = []
    for _ in range(2):
        dhm2 = torch.empty(N, F, device=cuda); db2 = torch.empty(F, device=cuda); dws2 = torch.empty(F, device=cuda)
        call("tsg_sag_conv_bwd_fused", ptr(dxo), ptr(inv), ptr(score), ptr(h), ptr(dsw), ptr(wsv), ptr(dhm2), ptr(db2),
             ptr(dws2), N, F, ptr(ws), wsb, stream_ptr())
        outs.append((dhm2, db2, dws2))
    assert torch.equal(outs[0][0], dhm) and torch.equal(outs[0][1], db)
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    ref = (h.double().t() @ dsw.double()).cpu().numpy()
    np.testing.assert_allclose(outs[0][2].cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * max(1.0, N ** 0.5))
