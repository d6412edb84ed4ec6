"""The CPU oracle against its own brute-force loop restatements (integer ops: exact)."""
import numpy as np
import pytest
import torch

from oracle import pyg_ref as R
from tsg import synth


def _small_batch(shape="PROTEINS", G=6, seed=3):
    c = synth.make_corpus(shape, G, seed=seed)
    b = synth.pack(c)
    return (torch.from_numpy(b["x"]), torch.from_numpy(b["edge_index"]), torch.from_numpy(b["batch"]),
            b["node_ptr"])


def test_synth_is_deterministic_and_connected():
    a = synth.make_corpus("DD", 5, seed=777)
    b = synth.make_corpus("DD", 5, seed=777)
    assert np.array_equal(a.row, b.row) and np.array_equal(a.col, b.col)
    for g in range(5):
        n = a.num_nodes(g)
        r = a.row[a.edge_ptr[g]:a.edge_ptr[g + 1]]; c = a.col[a.edge_ptr[g]:a.edge_ptr[g + 1]]
        assert set(r.tolist()) == set(range(n))                # no isolated node
        assert np.all(r != c)
        code = r * n + c
        assert np.all(np.diff(code) > 0)                        # lexicographic, no duplicates
        assert set((c * n + r).tolist()) == set(code.tolist())  # symmetric


def test_csr_matches_loops():
    x, ei, batch, _ = _small_batch()
    n = x.size(0)
    # add a couple of pre-existing self loops to exercise add_remaining_self_loops
    ei = torch.cat([ei, torch.tensor([[0, 5], [0, 5]])], dim=1)
    ei2, norm = R.gcn_norm(ei, None, n)
    for by in ("dst", "src"):
        a = R.csr_from_coo(ei2, norm, n, by)
        b = R.csr_from_coo_loops(ei2, norm, n, by)
        for u, v in zip(a, b):
            assert torch.equal(u, v)


def test_spmm_orders_agree_bitwise():
    x, ei, batch, _ = _small_batch()
    n = x.size(0)
    ei2, norm = R.gcn_norm(ei, None, n)
    h = torch.randn(n, 8, generator=torch.Generator().manual_seed(0))
    a = R.spmm_coo_edge_order(ei2, norm, h, n)
    b = R.spmm_loops(ei2, norm, h, n)
    rp, ci, v, _ = R.csr_from_coo(ei2, norm, n)
    c = R.spmm_csr_sequential(rp, ci, v, h)
    assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("ratio", [0.5, 0.8, 0.33])
def test_topk_matches_loops(ratio):
    x, ei, batch, _ = _small_batch(G=9)
    g = torch.Generator().manual_seed(1)
    score = torch.randn(x.size(0), generator=g)
    score[::7] = score[0]                 # ties
    score[3] = float("nan")
    score[5] = -0.0; score[6] = 0.0
    assert torch.equal(R.topk(score, ratio, batch), R.topk_loops(score, ratio, batch))


def test_filter_adj_matches_loops():
    x, ei, batch, _ = _small_batch(G=5)
    score = torch.randn(x.size(0), generator=torch.Generator().manual_seed(2))
    perm = R.topk(score, 0.5, batch)
    a, _ = R.filter_adj(ei, None, perm, x.size(0))
    assert torch.equal(a, R.filter_adj_loops(ei, perm, x.size(0)))


def test_readouts():
    x, ei, batch, _ = _small_batch(G=4)
    f = torch.randn(x.size(0), 5, generator=torch.Generator().manual_seed(4))
    mx = R.global_max_pool(f, batch); mn = R.global_mean_pool(f, batch)
    for g in range(4):
        seg = f[batch == g]
        assert torch.equal(mx[g], seg.max(0).values)
        assert torch.allclose(mn[g], seg.mean(0), atol=1e-6)


def test_pairwise_distance_matches_torch():
    a = torch.randn(7, 16); b = torch.randn(7, 16)
    assert torch.allclose(R.pairwise_distance(a, b), torch.nn.functional.pairwise_distance(a, b, 2), atol=1e-6)
    z = torch.zeros(1, 4)
    assert abs(float(R.pairwise_distance(z, z)) - 2e-6) < 1e-9
    ea, ep, en = torch.randn(5, 8), torch.randn(5, 8), torch.randn(5, 8)
    loss, dp, dn = R.triplet_margin_loss(ea, ep, en, 1.5)
    ref = torch.nn.MarginRankingLoss(margin=1.5)(dp, dn, torch.full_like(dp, -1.0))
    assert torch.allclose(loss, ref)


def test_sag_net_oracle_runs_and_backprops():
    x, ei, batch, ptr = _small_batch("PROTEINS", G=5)
    p = R.init_sag_params(x.size(1), 16, 8, seed=777)
    for v in p.values():
        v.requires_grad_(True)
    out = R.sag_net_forward(p, x, ei, batch, 0.5)
    assert out.shape == (5, 8)
    out.sum().backward()
    assert all(v.grad is not None for v in p.values())


def test_onehot_features_make_conv1_a_table_gather():
    """K3c / tsg_spmm_label_dot rest on: x = onehot(label)  =>  x @ W == W[label] bit for bit (every other term of the
    dot product is +-0 * w) and x^T @ dY == the segment sum of dY rows by label.  Checked on the oracle's own GCNConv."""
    x, ei, batch, _ = _small_batch("DD", 4, seed=9)
    label = x.argmax(1)
    assert torch.equal(x, torch.nn.functional.one_hot(label, x.size(1)).float())
    g = torch.Generator().manual_seed(0)
    W = torch.randn(x.size(1), 32, generator=g); b = torch.randn(32, generator=g)
    assert torch.equal(x @ W, W[label])
    n = x.size(0)
    ei2, norm = R.gcn_norm(ei, None, n)
    ref = R.gcn_conv(x, ei, W, b)
    via_table = R.spmm_coo_edge_order(ei2, norm, W[label], n) + b
    assert torch.equal(ref, via_table)
    dy = torch.randn(n, 32, generator=g)
    seg = torch.zeros(x.size(1), 32, dtype=torch.float64).index_add_(0, label, dy.double())
    np.testing.assert_allclose((x.double().t() @ dy.double()).numpy(), seg.numpy(), rtol=0, atol=1e-12)


def test_coalesced_edge_list_is_its_own_csr_in_both_orientations():
    """K1b's fast path rests on: for a list sorted by (row, col), loop free, without duplicates and symmetric, the
    dst-major and the src-major CSR of the self-loop-augmented operator are the SAME arrays -- the list itself with
    one loop closing every row -- and the edge id of the k-th dst-major entry of row r is the id of the reverse of
    the k-th edge of run r.  Checked against the oracle's stable counting sort (the K1 contract)."""
    c = synth.make_corpus("DD", 3, seed=5)
    for gidx in range(3):
        n = c.num_nodes(gidx)
        sl = slice(int(c.edge_ptr[gidx]), int(c.edge_ptr[gidx + 1]))
        ei = torch.from_numpy(np.stack([c.row[sl], c.col[sl]]).astype(np.int64))
        rp_d, ci_d, v_d, eid_d = R.gcn_csr(ei, n, by="dst")
        rp_s, ci_s, v_s, eid_s = R.gcn_csr(ei, n, by="src")
        assert torch.equal(rp_d, rp_s) and torch.equal(ci_d, ci_s) and torch.equal(v_d, v_s)
        m = ei.size(1)
        # the list itself: entry of edge e = (r, c) sits at e + r, the loop of row r at rp[r + 1] - 1
        pos = torch.arange(m) + ei[0]
        assert torch.equal(ci_s[pos].long(), ei[1]) and torch.equal(eid_s[pos].long(), torch.arange(m))
        code = (ei[0] * n + ei[1]).tolist()
        rev = torch.tensor([code.index(int(cc) * n + int(rr)) for rr, cc in zip(ei[0], ei[1])])
        assert torch.equal(eid_d[pos].long(), rev)
        loops = rp_s[1:].long() - 1
        assert torch.equal(ci_s[loops].long(), torch.arange(n))
