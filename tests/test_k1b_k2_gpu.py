"""K1b (per-graph shared-memory CSR build) and the chunk-pipelined K2 variants: bit-exact against the
generic kernels and the CPU oracle (index work and index_add_-order sums are integer-exact claims)."""
import os

import numpy as np
import pytest
import torch

from oracle import pyg_ref as R
from tsg import synth

pytestmark = pytest.mark.gpu


def _batch(shape, G, seed=11):
    c = synth.make_corpus(shape, G, seed=seed)
    b = synth.pack(c)
    return torch.from_numpy(b["edge_index"]), b["node_ptr"]


def _check(y, ref, exact):
    """exact mode is bit-identical to the index_add_-order oracle; the default keeps the order but fuses
    the product rounding: <= 1 ulp per term."""
    if exact:
        assert torch.equal(y, ref)
    else:
        assert float((y - ref).abs().max()) <= 2e-6 * max(float(ref.abs().max()), 1e-30)


def _same_csr(a, b):
    nnz = int(a.rowptr[-1])
    assert int(b.rowptr[-1]) == nnz
    for x, y in ((a.rowptr, b.rowptr), (a.t_rowptr, b.t_rowptr)):
        assert torch.equal(x, y)
    for x, y in ((a.colidx, b.colidx), (a.t_colidx, b.t_colidx), (a.eid, b.eid), (a.t_eid, b.t_eid)):
        assert torch.equal(x[:nnz], y[:nnz])
    for x, y in ((a.val, b.val), (a.t_val, b.t_val)):
        assert torch.equal(x[:nnz].view(torch.int32), y[:nnz].view(torch.int32))


@pytest.mark.parametrize("shape,G", [("PROTEINS", 33), ("DD", 9), ("JANY", 4)])
def test_k1b_equals_k1_and_oracle(cuda, shape, G):
    from tsg import ops
    ei, nptr = _batch(shape, G)
    n = int(nptr[-1])
    # pre-existing self loops inside graph 0 and graph 1 (kept contiguous with their graphs)
    e0 = int((ei[0] < nptr[1]).sum())
    loops = torch.tensor([[0, 2, 2], [0, 2, 2]])
    ei = torch.cat([ei[:, :e0], loops, ei[:, e0:]], dim=1)
    el = ops.EdgeList.from_edge_index(ei.to(cuda))
    a = ops.build_csr(el, n, want_eid=True)
    b = ops.build_csr_graphs(el, torch.from_numpy(nptr).to(cuda), n, int(np.diff(nptr).max()), want_eid=True)
    _same_csr(a, b)
    ei2, norm = R.gcn_norm(ei, None, n)
    o_rp, o_ci, o_v, _ = R.csr_from_coo(ei2, norm, n, "dst")
    nnz = int(o_rp[-1])
    assert torch.equal(b.rowptr.cpu(), o_rp) and torch.equal(b.colidx.cpu()[:nnz], o_ci)
    assert torch.equal(b.val.cpu()[:nnz].view(torch.int32), o_v.view(torch.int32))


def test_k1b_device_count_and_empty_graphs(cuda):
    """capacity buffer longer than the valid prefix (filter_adj output) + graphs that lost every edge."""
    from tsg import ops
    ei, nptr = _batch("PROTEINS", 12)
    n, E = int(nptr[-1]), ei.size(1)
    # drop all edges of graphs 3 and 7 (edge-free graphs still get their self loops)
    g_of_e = np.searchsorted(nptr, ei[0].numpy(), side="right") - 1
    keep = torch.from_numpy(~np.isin(g_of_e, [3, 7]))
    ei = ei[:, keep]
    E2 = ei.size(1)
    pad = torch.cat([ei, torch.full((2, E - E2 + 5), 1, dtype=torch.int64)], dim=1).to(cuda)
    el = ops.EdgeList(pad[0], pad[1], pad.size(1), torch.tensor([E2], device=cuda))
    a = ops.build_csr(el, n, want_eid=True)
    b = ops.build_csr_graphs(el, torch.from_numpy(nptr).to(cuda), n, int(np.diff(nptr).max()), want_eid=True)
    _same_csr(a, b)


def test_k1b_single_node_graphs(cuda):
    from tsg import ops
    nptr = np.arange(0, 41, dtype=np.int64)                 # 40 graphs of one node, no edges
    el = ops.EdgeList(torch.zeros(1, dtype=torch.int64, device=cuda), torch.zeros(1, dtype=torch.int64, device=cuda), 0)
    b = ops.build_csr_graphs(el, torch.from_numpy(nptr).to(cuda), 40, 1)
    assert torch.equal(b.rowptr.cpu(), torch.arange(41, dtype=torch.int32))
    assert torch.equal(b.colidx.cpu()[:40], torch.arange(40, dtype=torch.int32))
    assert torch.equal(b.val.cpu()[:40], torch.ones(40))


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("F,G", [(4, 5), (8, 5), (24, 5), (32, 5), (32, 320), (64, 200), (128, 90), (160, 5)])
def test_spmm_kernels(cuda, exact, F, G):
    """small grids (256-thread CTAs) and large ones (1024-thread CTAs + TMA L2 prefetch), every lane
    layout, exact and fused accumulation, val = NULL, bias + ReLU epilogue."""
    from tsg import ops
    ei, nptr = _batch("DD", G)
    n = int(nptr[-1])
    g = torch.Generator().manual_seed(F)
    h = torch.randn(n, F, generator=g)
    bias = torch.randn(F, generator=g)
    csr = ops.build_csr(ops.EdgeList.from_edge_index(ei.to(cuda)), n)
    ei2, norm = R.gcn_norm(ei, None, n)
    ref = R.spmm_coo_edge_order(ei2, norm, h, n)
    y = ops.spmm_raw(csr.rowptr, csr.colidx, csr.val, h.to(cuda), exact=exact)
    _check(y.cpu(), ref, exact)
    y2 = ops.spmm_raw(csr.rowptr, csr.colidx, csr.val, h.to(cuda), bias.to(cuda), relu=True, exact=exact)
    _check(y2.cpu(), torch.relu(ref + bias), exact)
    yt = ops.spmm_raw(csr.t_rowptr, csr.t_colidx, None, h.to(cuda), exact=exact)       # val = NULL => weights 1
    ref_t = torch.zeros(n, F).index_add_(0, ei2[0], h[ei2[1]])
    _check(yt.cpu(), ref_t, exact)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("n", [1000 + 13, 160_000 + 7])
def test_spmm_hub_rows_and_empty_rows(cuda, exact, n):
    """a hub row with hundreds of entries, rows without entries (RAW mode), a row count that is not a
    multiple of the CTA tile, at a small and a large (1024-thread path) size."""
    from tsg import ops
    rng = np.random.default_rng(3)
    m = 6 * n
    src = rng.integers(0, n, m); dst = rng.integers(0, n, m)
    hub_src = rng.integers(0, n, 700); hub_dst = np.full(700, 40)          # row 40 has ~700 entries
    row = np.concatenate([src, hub_src]); col = np.concatenate([dst, hub_dst])
    keep = col % 7 != 3                                                   # rows = 3 mod 7 are empty
    ei = torch.from_numpy(np.stack([row[keep], col[keep]]).astype(np.int64))
    w = torch.rand(ei.size(1), generator=torch.Generator().manual_seed(1))
    csr = ops.build_csr(ops.EdgeList.from_edge_index(ei.to(cuda)), n, mode=ops.CSR_RAW, edge_weight=w.to(cuda))
    h = torch.randn(n, 32, generator=torch.Generator().manual_seed(2))
    ref = R.spmm_coo_edge_order(ei, w, h, n)
    y = ops.spmm_raw(csr.rowptr, csr.colidx, csr.val, h.to(cuda), exact=exact)
    _check(y.cpu(), ref, exact)


@pytest.mark.parametrize("shape,G,ratio", [("DD", 9, 0.5), ("PROTEINS", 33, 0.5), ("JANY", 4, 0.25), ("DD", 5, 1.0)])
def test_k1c_csr_filter_equals_rebuild(cuda, shape, G, ratio):
    """K1c(CSR, perm, inv) == K1b(filter_adj(edges, perm)): rowptr / colidx / val of both orientations, bit for bit."""
    from tsg import ops
    from tsg._lib import call, lib, ptr, stream_ptr, workspace
    ei, nptr = _batch(shape, G)
    n = int(nptr[-1])
    batch = torch.from_numpy(np.repeat(np.arange(G), np.diff(nptr)))
    score = torch.randn(n, generator=torch.Generator().manual_seed(G))
    perm = R.topk(score, ratio, batch).to(cuda)
    k = perm.numel()
    el = ops.EdgeList.from_edge_index(ei.to(cuda))
    gptr = torch.from_numpy(nptr).to(cuda)
    old = ops.build_csr_graphs(el, gptr, n, int(np.diff(nptr).max()))
    el2, inv = ops.filter_adj(el, perm, n)
    kptr = ops.topk_sizes(gptr, ratio)
    ref = ops.build_csr_graphs(el2, kptr, k, int(np.diff(nptr).max()))
    inv2 = torch.empty(n, dtype=torch.int32, device=cuda)
    call("tsg_inv_perm", ptr(perm), k, n, ptr(inv2), stream_ptr())
    assert torch.equal(inv, inv2)
    cap = int(old.rowptr[-1])
    i32 = dict(dtype=torch.int32, device=cuda)
    out = [torch.empty(k + 1, **i32), torch.empty(cap, **i32), torch.empty(cap, dtype=torch.float32, device=cuda),
           torch.empty(k + 1, **i32), torch.empty(cap, **i32), torch.empty(cap, dtype=torch.float32, device=cuda)]
    wsb = lib.tsg_csr_filter_workspace_bytes(k)
    ws = workspace(wsb, cuda)
    call("tsg_csr_filter", ptr(old.rowptr), ptr(old.colidx), ptr(old.t_rowptr), ptr(old.t_colidx), ptr(perm), ptr(inv2), k,
         *[ptr(t) for t in out], ptr(ws), wsb, stream_ptr())
    nnz = int(ref.rowptr[-1])
    assert torch.equal(out[0], ref.rowptr) and torch.equal(out[3], ref.t_rowptr)
    assert torch.equal(out[1][:nnz], ref.colidx[:nnz]) and torch.equal(out[4][:nnz], ref.t_colidx[:nnz])
    assert torch.equal(out[2][:nnz].view(torch.int32), ref.val[:nnz].view(torch.int32))
    assert torch.equal(out[5][:nnz].view(torch.int32), ref.t_val[:nnz].view(torch.int32))


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("shape,G,F", [("DD", 40, 32), ("DD", 300, 32), ("PROTEINS", 500, 16), ("JANY", 30, 8), ("DD", 9, 24)])
def test_spmm_tma_equals_spmm(cuda, exact, shape, G, F):
    """TMA-staged tile kernel == k_spmm_g bit for bit (same order, same rounding), forward and transposed, incl.
    graphs that exceed the stage capacity (DD graphs > 544 rows take the in-kernel global path)."""
    from tsg import ops
    ei, nptr = _batch(shape, G)
    n = int(nptr[-1])
    el = ops.EdgeList.from_edge_index(ei.to(cuda))
    gptr = torch.from_numpy(nptr).to(cuda)
    csr = ops.build_csr_graphs(el, gptr, n, int(np.diff(nptr).max()))
    g = torch.Generator().manual_seed(F + G)
    h = torch.randn(n, F, generator=g).to(cuda)
    bias = torch.randn(F, generator=g).to(cuda)
    for rp, ci, v in ((csr.rowptr, csr.colidx, csr.val), (csr.t_rowptr, csr.t_colidx, csr.t_val), (csr.rowptr, csr.colidx, None)):
        ref = ops.spmm_raw(rp, ci, v, h, bias, relu=True, exact=exact)
        got = ops.spmm_raw(rp, ci, v, h, bias, relu=True, exact=exact, tile_ptr=gptr, tma=True)
        assert torch.equal(got, ref)


@pytest.mark.parametrize("F", [32, 16, 128, 24, 6, 256])
def test_spmm_dot_epilogue_is_the_k3_product(cuda, F):
    """tsg_spmm_dot: Y bit-identical to tsg_spmm, dot_out bit-identical to tsg_linear_fwd(Y, w, out_feat=1) -- from K2's
    epilogue when F % 4 == 0 and F <= 128, from the K3 call behind it otherwise (F = 6, 256)."""
    from tsg import ops, synth
    from tsg._lib import call, ptr, stream_ptr
    c = synth.make_corpus("DD", 700, seed=21)          # enough rows for the 1024-thread configuration
    b = synth.pack(c)
    n = int(c.node_ptr[-1])
    csr = ops.build_csr(ops.EdgeList.from_edge_index(torch.from_numpy(b["edge_index"]).to(cuda)), n)
    g = torch.Generator().manual_seed(F)
    H = torch.randn(n, F, generator=g).to(cuda); bias = torch.randn(F, generator=g).to(cuda)
    w = torch.randn(F, 1, generator=g).to(cuda)
    for rows in (n, 300):                                # big (1024-thread) and small (256-thread) launch shapes
        Y0 = ops.spmm_raw(csr.rowptr[:rows + 1], csr.colidx, csr.val, H, bias, relu=True)
        d0 = ops.linear_raw(Y0, w, None, False)
        Y1 = torch.empty(rows, F, device=cuda); d1 = torch.empty(rows, device=cuda)
        call("tsg_spmm_dot", ptr(csr.rowptr), ptr(csr.colidx), ptr(csr.val), ptr(H), ptr(bias), ptr(Y1), ptr(w), ptr(d1),
             rows, F, ops.SPMM_RELU, stream_ptr())
        assert torch.equal(Y0, Y1)
        assert torch.equal(d0.view(-1), d1)
