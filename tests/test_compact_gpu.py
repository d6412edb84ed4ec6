"""Compact level-0 input (labels + graph-local int32 endpoints) == the expanded PyG wire format.

K3c: onehot(label) @ W as a row gather (bit-identical to the dense K3 product) and its dW as a fixed-order segment sum
(fp32 summation-order tolerance against an fp64 reference, deterministic run to run); K1b on local endpoints
(bit-identical CSR, both launch size classes); the executor's compact entries against the dense ones."""
import numpy as np
import pytest
import torch

from tsg import synth

pytestmark = pytest.mark.gpu


def _compact(c, dev):
    from tsg import ops
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    return ops.CompactBatch(t(c.node_label, np.int32), t(c.row, np.int32), t(c.col, np.int32),
                            t(c.node_ptr, np.int64), t(c.edge_ptr, np.int64), c.num_node_labels)


@pytest.mark.parametrize("N,K,M", [(5000, 89, 32), (777, 3, 128), (1000, 7, 5), (1, 4, 32), (0, 4, 32)])
def test_embed_fwd_is_the_dense_product(cuda, N, K, M):
    from tsg import ops
    g = torch.Generator().manual_seed(N + K)
    label = torch.randint(0, K, (N,), generator=g, dtype=torch.int32)
    if N > 10:
        label[3], label[7] = -1, K          # outside [0, K): the all-zero one-hot row
    W = torch.randn(K, M, generator=g)
    x = torch.zeros(N, K)
    ok = (label >= 0) & (label < K)
    x[torch.arange(N)[ok], label[ok].long()] = 1.0
    out = ops.embed_fwd(W.to(cuda), label.to(cuda))
    if N > 0:
        ref = ops.linear(x.to(cuda), W.to(cuda), None)
        assert torch.equal(out, ref)
        assert torch.equal(out.cpu()[ok], W[label[ok].long()])
    assert out.shape == (N, M)


@pytest.mark.parametrize("N,K,M", [(200000, 89, 32), (5000, 3, 128), (999, 7, 5), (40, 600, 64), (0, 4, 8)])
def test_embed_bwd_weight(cuda, N, K, M):
    from tsg import ops
    g = torch.Generator().manual_seed(N + M)
    label = torch.randint(0, K, (N,), generator=g, dtype=torch.int32)
    if N > 10:
        label[5], label[9] = -3, K + 2
    dy = torch.randn(N, M, generator=g)
    ok = (label >= 0) & (label < K)
    ref = torch.zeros(K, M, dtype=torch.float64)
    ref.index_add_(0, label[ok].long(), dy[ok].double())
    got = ops.embed_bwd_weight(label.to(cuda), dy.to(cuda), K)
    again = ops.embed_bwd_weight(label.to(cuda), dy.to(cuda), K)
    assert torch.equal(got, again)                                     # fixed summation order
    # tolerance: fp32 accumulation of ~N/K terms of unit variance
    np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5 * max(1.0, (N / K) ** 0.5))


def test_embed_bwd_weight_rejects_oversize_table(cuda):
    from tsg import _lib
    assert _lib.lib.tsg_embed_bwd_weight_workspace_bytes(89, 32) > 0
    assert _lib.lib.tsg_embed_bwd_weight_workspace_bytes(4096, 1024) == 0


def _big_graph_corpus():
    """a few DD graphs + one 1,500-node graph (second launch size class) + a star with a 300-edge hub row (rows longer
    than the per-thread sort limit) + an edge-free and a one-node graph"""
    base = synth.make_corpus("DD", 6, seed=11)
    rng = np.random.default_rng(0)
    n_big, m_big = 1500, 4000
    r = rng.integers(0, n_big, m_big); c = rng.integers(0, n_big, m_big)      # includes a few self loops
    leaves = rng.permutation(np.arange(1, 301)); hub = np.zeros(300, np.int64)
    star_r = np.concatenate([hub, leaves, [7, 7]]); star_c = np.concatenate([leaves, hub, [7, 9]])   # + a loop, + a duplicate-free extra
    z = np.zeros(0, np.int64)
    rows = [base.row, np.concatenate([r, c]), star_r, z, z]
    cols = [base.col, np.concatenate([c, r]), star_c, z, z]
    sizes = [n_big, 301, 5, 1]
    edges = [2 * m_big, star_r.shape[0], 0, 0]
    node_ptr = np.concatenate([base.node_ptr, base.node_ptr[-1] + np.cumsum(sizes)]).astype(np.int64)
    edge_ptr = np.concatenate([base.edge_ptr, base.edge_ptr[-1] + np.cumsum(edges)]).astype(np.int64)
    labels = np.concatenate([base.node_label, rng.integers(0, base.num_node_labels, sum(sizes)).astype(np.int32)])
    y = np.concatenate([base.y, [0, 1, 0, 1]])
    return synth.Corpus("DD", node_ptr, edge_ptr, np.concatenate(rows), np.concatenate(cols), labels, y, base.num_node_labels)


def test_csr_from_local_endpoints(cuda):
    from tsg import ops
    c = _big_graph_corpus()
    cb = _compact(c, cuda)
    b = synth.pack(c)
    ei = torch.from_numpy(b["edge_index"]).to(cuda)
    n = int(c.node_ptr[-1]); mx = int(np.diff(c.node_ptr).max())
    assert mx > 1024
    a = ops.build_csr_graphs(ops.EdgeList.from_edge_index(ei), cb.node_ptr, n, mx)
    l = ops.build_csr_graphs_local(cb, n, mx)
    nnz = int(l.rowptr[-1])                  # input self loops are dropped: the arrays' tail past nnz is unused
    assert nnz < l.colidx.numel() and nnz == int(a.rowptr[-1]) == int(l.t_rowptr[-1])
    g = ops.build_csr(ops.EdgeList.from_edge_index(ei), n)                  # generic K1
    for other in (a, g):
        for name in ("rowptr", "t_rowptr"):
            assert torch.equal(getattr(other, name), getattr(l, name)), name
        for name in ("colidx", "val", "t_colidx", "t_val"):
            assert torch.equal(getattr(other, name)[:nnz], getattr(l, name)[:nnz]), name
    x, ei2 = cb.expand()
    assert torch.equal(ei2, ei) and torch.equal(x, torch.from_numpy(b["x"]).to(cuda))


@pytest.mark.parametrize("shape,G,nhid", [("DD", 12, 32), ("PROTEINS", 50, 32), ("DD", 6, 128), ("big", 0, 32)])
def test_compact_encoder_matches_dense(cuda, shape, G, nhid):
    from tsg import nn as tnn
    c = _big_graph_corpus() if shape == "big" else synth.make_corpus(shape, G, seed=5)
    G = c.num_graphs
    b = synth.pack(c)
    x = torch.from_numpy(b["x"]).to(cuda)
    ei = torch.from_numpy(b["edge_index"]).to(cuda)
    cb = _compact(c, cuda)
    torch.manual_seed(3)
    model = tnn.PackedSAGNet(c.num_node_labels, nhid, 8, 0.5, 0.0).to(cuda)
    cot = torch.randn(G, 8, generator=torch.Generator().manual_seed(1)).to(cuda)
    res = []
    for inp in ((x, ei), (cb, None)):
        model.zero_grad(set_to_none=True)
        out = model(inp[0], inp[1], b["node_ptr"])
        (out * cot).sum().backward()
        res.append((out.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0])
    for k in res[0][1]:
        if k == "conv1.weight":
            scale = float(res[0][1][k].abs().max()) + 1e-12
            assert float((res[0][1][k] - res[1][1][k]).abs().max()) <= 1e-5 * scale, k
        else:
            assert torch.equal(res[0][1][k], res[1][1][k]), k


def test_compact_falls_back_when_executor_is_off(cuda):
    from tsg import nn as tnn
    c = synth.make_corpus("PROTEINS", 10, seed=2)
    b = synth.pack(c)
    cb = _compact(c, cuda)
    torch.manual_seed(0)
    model = tnn.PackedSAGNet(c.num_node_labels, 16, 4, 0.5, 0.0).to(cuda)
    a = model(cb, None, b["node_ptr"]).detach().clone()
    tnn.USE_EXECUTOR = False
    try:
        o = model(cb, None, b["node_ptr"]).detach().clone()          # expand() + op-by-op
    finally:
        tnn.USE_EXECUTOR = True
    assert torch.equal(a, o)
    with pytest.raises(ValueError):
        tnn.PackedSAGNet(c.num_node_labels + 1, 16, 4, 0.5, 0.0).to(cuda)(cb, None, b["node_ptr"])


def test_resident_corpus_compact_gather(cuda):
    """DeviceCorpus.pack_compact (K0 compact gather) expands to exactly what DeviceCorpus.pack writes, repeated ids
    included; step_from_ids takes the compact route for a model that accepts it."""
    from tsg.feeder import DeviceCorpus
    c = synth.make_corpus("PROTEINS", 30, seed=9)
    dc = DeviceCorpus(c, cuda)
    ids = np.array([4, 4, 0, 29, 17, 4], dtype=np.int64)
    x, ei, nptr = dc.pack(ids)
    cb, nptr2 = dc.pack_compact(ids)
    assert np.array_equal(nptr, nptr2)
    x2, ei2 = cb.expand()
    assert torch.equal(x, x2) and torch.equal(ei, ei2)
    sel = synth.select(c, ids)
    assert np.array_equal(cb.label.cpu().numpy(), sel.node_label)
    assert np.array_equal(cb.row.cpu().numpy(), sel.row) and np.array_equal(cb.col.cpu().numpy(), sel.col)
    dense = DeviceCorpus(c, cuda, dense_x=np.random.default_rng(0).random((int(c.node_ptr[-1]), 5), dtype=np.float32))
    with pytest.raises(RuntimeError):
        dense.pack_compact(ids)


@pytest.mark.parametrize("F,with_dot", [(32, True), (16, True), (128, True), (64, False)])
def test_label_table_spmm_is_embed_then_spmm(cuda, F, with_dot):
    """tsg_spmm_label_dot (W[label] gathered inside K2) == tsg_embed_fwd -> tsg_spmm_dot, bit for bit, in both launch
    shapes; labels outside the table contribute the zero row."""
    from tsg import ops
    from tsg._lib import call, ptr, stream_ptr
    c = synth.make_corpus("DD", 700, seed=4)
    b = synth.pack(c)
    n = int(c.node_ptr[-1]); K = c.num_node_labels
    csr = ops.build_csr(ops.EdgeList.from_edge_index(torch.from_numpy(b["edge_index"]).to(cuda)), n)
    g = torch.Generator().manual_seed(F)
    label = torch.from_numpy(c.node_label.astype(np.int32)).clone()
    label[5], label[77], label[n - 1] = -1, K, K + 9
    label = label.to(cuda)
    W = torch.randn(K, F, generator=g).to(cuda); bias = torch.randn(F, generator=g).to(cuda)
    w = torch.randn(F, 1, generator=g).to(cuda)
    xw = ops.embed_fwd(W, label)
    for rows in (n, 300):
        Y0 = torch.empty(rows, F, device=cuda); d0 = torch.empty(rows, device=cuda)
        call("tsg_spmm_dot", ptr(csr.rowptr), ptr(csr.colidx), ptr(csr.val), ptr(xw), ptr(bias), ptr(Y0), ptr(w), ptr(d0),
             rows, F, ops.SPMM_RELU, stream_ptr())
        Y1 = torch.empty(rows, F, device=cuda); d1 = torch.empty(rows, device=cuda)
        call("tsg_spmm_label_dot", ptr(csr.rowptr), ptr(csr.colidx), ptr(csr.val), ptr(W), ptr(label), K, ptr(bias), ptr(Y1),
             ptr(w) if with_dot else None, ptr(d1) if with_dot else None, rows, F, ops.SPMM_RELU, stream_ptr())
        assert torch.equal(Y0, Y1)
        if with_dot:
            assert torch.equal(d0, d1)
