"""CUDA SAGPool path vs (1) the fixture computed by the UNMODIFIED reference glue (Code/sag/network.py Net +
layers.py SAGPool over the oracle operators: tests/golden/sag_glue.npz) and (2) the oracle on bench-shape batches,
for all three ways the product runs the encoder: operator by operator, the K10 executor on the PyG wire format, and
the executor's compact entries (what bench.py's `e2e` / `value_compact_input` time).  Permutations are asserted level
by level on every path (the executor's are read back from its arena: tsg_sag_arena_locate); the pooled levels'
edge lists are asserted as the next level's CSR (the executor never materialises filter_adj's output: K1c)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from golden_util import load_sag_glue
from oracle import pyg_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _compact(d, dev):
    from tsg import ops
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev)
    return ops.CompactBatch(t(d["label"], torch.int32), t(d["local_row"], torch.int32), t(d["local_col"], torch.int32),
                            t(d["node_ptr"], torch.int64), t(d["edge_ptr"], torch.int64), d["F"])


def _run(model, d, dev, path):
    """-> (emb, perms[3], csr_next[2] = (rowptr, colidx, val) of levels 1, 2, or filtered edge lists)."""
    from tsg import nn as tnn
    x, ei = d["x"].to(dev), d["edge_index"].to(dev)
    if path == "ops":
        emb, aux = model(x, ei, d["node_ptr"], return_aux=True)
        return emb, [p.cpu() for p in aux["perm"]], [e.edge_index().cpu() for e in aux["edges"]], None
    tnn.KEEP_ARENA = True
    try:
        emb = model(_compact(d, dev), None, d["node_ptr"]) if path == "compact" else model(x, ei, d["node_ptr"])
        shape, arena = tnn.LAST_ARENA
    finally:
        tnn.KEEP_ARENA = False
    perms = [tnn.sag_arena_view(shape, arena, l, "perm").cpu() for l in range(3)]
    csrs = [tuple(tnn.sag_arena_view(shape, arena, l, f).cpu() for f in ("rowptr", "colidx", "val")) for l in range(3)]
    return emb, perms, None, csrs


def _check_against(d, emb, perms, edges, csrs, grads, ref_perm, ref_ei, ref_emb, ref_grads, n_levels):
    for lvl in range(3):
        assert torch.equal(perms[lvl], ref_perm[lvl]), f"perm level {lvl}"
    if edges is not None:
        for lvl in range(3):
            assert torch.equal(edges[lvl], ref_ei[lvl]), f"filter_adj level {lvl}"
    else:
        # level l+1's operator = gcn_norm(filter_adj(level l)) in CSR: integer structure bit-exact, values to 1 ulp
        # (K1c renormalises from degrees; exactness of K1c vs K1b(filter_adj) is tests/test_k1b_k2_gpu.py's)
        for lvl in (1, 2):
            n = int(n_levels[lvl][-1])
            rp, ci, v, _ = R.gcn_csr(ref_ei[lvl - 1], n)
            nnz = int(rp[-1])
            assert torch.equal(csrs[lvl][0], rp), f"CSR rowptr level {lvl}"
            assert torch.equal(csrs[lvl][1][:nnz], ci), f"CSR colidx level {lvl}"
            assert torch.equal(csrs[lvl][2][:nnz], v), f"CSR val level {lvl}"
    assert rel_err(emb, ref_emb) <= TOL
    for k, g in ref_grads.items():
        assert rel_err(grads[k], g) <= TOL, k


@pytest.mark.parametrize("path", ["ops", "executor", "compact"])
def test_cuda_paths_match_the_reference_glue_fixture(cuda, path):
    from tsg import nn as tnn
    d = load_sag_glue()
    model = tnn.PackedSAGNet(d["F"], d["nhid"], d["C"], 0.5, 0.5).to(cuda)
    model.load_state_dict(d["params"])
    model.eval()
    emb, perms, edges, csrs = _run(model, d, cuda, path)
    (emb * d["cot"].to(cuda)).sum().backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    _check_against(d, emb, perms, edges, csrs, grads, d["perm"], d["ei_lvl"], d["emb"], d["grads"], d["level_ptr"])


@pytest.mark.parametrize("path", ["executor", "compact"])
@pytest.mark.parametrize("shape,G,nhid", [("DD", 24, 32), ("DD", 5, 128), ("PROTEINS", 40, 32)])
def test_executor_vs_oracle_direct(cuda, path, shape, G, nhid):
    """The bench path against the ORACLE (not against the op-by-op path): perms per level, next-level CSRs,
    embeddings and every parameter gradient, on batches of the bench's dataset shape."""
    from tsg import nn as tnn, synth
    from tsg.nn import host_level_ptrs
    C = 16
    corpus = synth.make_corpus(shape, G, seed=101)
    if shape == "PROTEINS":          # 3 node labels make exact score ties: widen the alphabet
        rng = np.random.default_rng(1)
        corpus.node_label[:] = rng.integers(0, 89, corpus.node_label.shape[0]); corpus.num_node_labels = 89
    b = synth.pack(corpus)
    d = dict(F=corpus.num_node_labels, label=torch.from_numpy(corpus.node_label), x=torch.from_numpy(b["x"]),
             edge_index=torch.from_numpy(b["edge_index"]), node_ptr=b["node_ptr"],
             local_row=torch.from_numpy(corpus.row), local_col=torch.from_numpy(corpus.col), edge_ptr=corpus.edge_ptr)
    params = R.init_sag_params(d["F"], nhid, C, seed=12)
    model = tnn.PackedSAGNet(d["F"], nhid, C, 0.5, 0.5).to(cuda)
    model.load_state_dict(params)
    model.eval()
    cot = torch.randn(G, C, generator=torch.Generator().manual_seed(2))
    emb, perms, edges, csrs = _run(model, d, cuda, path)
    (emb * cot.to(cuda)).sum().backward()
    po = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    emb_o, aux = R.sag_net_forward(po, d["x"], d["edge_index"], torch.from_numpy(b["batch"]), 0.5, return_aux=True)
    (emb_o * cot).sum().backward()
    _check_against(d, emb, perms, edges, csrs, {k: p.grad for k, p in model.named_parameters()}, aux["perm"],
                   aux["edge_index"], emb_o, {k: v.grad for k, v in po.items()}, host_level_ptrs(b["node_ptr"], 0.5))


def test_device_packers_match_the_oracle_collation(cuda):
    """K0 (tsg_pack_batch from an HBM-resident corpus) and the compact form's expand() against
    oracle.pyg_ref.batch_from_data_list (PyG Batch.from_data_list), not against each other."""
    from tsg import synth
    from tsg.feeder import DeviceCorpus
    corpus = synth.make_corpus("DD", 20, seed=8)
    ids = np.array([3, 3, 19, 0, 7, 12, 1], np.int64)
    xs, eis, ys = [], [], []
    for g in ids:
        one = synth.pack(corpus, [int(g)])
        xs.append(torch.from_numpy(one["x"])); eis.append(torch.from_numpy(one["edge_index"])); ys.append(torch.from_numpy(one["y"]))
    x_o, ei_o, batch_o, _ = R.batch_from_data_list(xs, eis, ys)
    dc = DeviceCorpus(corpus, cuda)
    x, ei, nptr = dc.pack(ids)
    assert torch.equal(x.cpu(), x_o) and torch.equal(ei.cpu(), ei_o)
    assert np.array_equal(np.repeat(np.arange(len(ids)), np.diff(nptr)), batch_o.numpy())
    cb, nptr2 = dc.pack_compact(ids)
    x2, ei2 = cb.expand()
    assert torch.equal(x2.cpu(), x_o) and torch.equal(ei2.cpu(), ei_o) and np.array_equal(nptr, nptr2)
