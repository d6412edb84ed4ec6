"""K13 graph-resident SAGPool forward (one CTA carries a graph through all three levels in shared memory; the
forward-only entry tsg_sag_encoder_embed_compact that torch.no_grad() forwards take) against the kernel-per-operator
executor (tsg_sag_set_fused(0)), which is itself pinned to the oracle and to the reference glue: embeddings, per-level
perm, score and h bit-identical."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from tsg import synth

pytestmark = pytest.mark.gpu


def _compact(corpus, dev):
    from tsg import ops
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
    return ops.CompactBatch(t(corpus.node_label, np.int32), t(corpus.row, np.int32), t(corpus.col, np.int32),
                            t(corpus.node_ptr, np.int64), t(corpus.edge_ptr, np.int64), corpus.num_node_labels,
                            int(np.diff(corpus.edge_ptr).max()))


def _embed(model, cb, node_ptr, fused):
    """forward under torch.no_grad() (= tsg_sag_encoder_embed_compact) with the graph-resident kernels on / off."""
    from tsg import _lib, nn as tnn
    prev = _lib.lib.tsg_sag_set_fused(int(fused))
    tnn.KEEP_ARENA = True
    try:
        before = _lib.kernel_launches
        with torch.no_grad():
            emb = model(cb, None, node_ptr)
        shape, arena = tnn.LAST_ARENA
        saved = {f"{f}{l}": tnn.sag_arena_view(shape, arena, l, f).clone() for l in range(3) for f in ("perm", "score", "h")}
        tnn.check_fused_status()
    finally:
        tnn.KEEP_ARENA = False
        _lib.lib.tsg_sag_set_fused(prev)
    return emb.clone(), saved


@pytest.mark.parametrize("shape,G,nhid", [("DD", 40, 32), ("PROTEINS", 300, 32), ("JANY", 12, 32), ("DD", 9, 64), ("PROTEINS", 64, 128),
                                          ("DD", 700, 32)])
def test_fused_forward_is_bit_identical(cuda, shape, G, nhid):
    from tsg import nn as tnn
    c = synth.make_corpus(shape, G, seed=17)
    cb = _compact(c, cuda)
    torch.manual_seed(5)
    model = tnn.PackedSAGNet(c.num_node_labels, nhid, 16, 0.5, 0.0).to(cuda)
    with torch.no_grad():                         # non-zero conv / score biases
        for k, p in model.named_parameters():
            if k.endswith("bias") and (k.startswith("conv") or k.startswith("pool")):
                p.copy_(torch.randn_like(p) * 0.1)
    model.eval()
    e0, s0 = _embed(model, cb, c.node_ptr, fused=False)
    e1, s1 = _embed(model, cb, c.node_ptr, fused=True)
    for l in range(3):
        assert torch.equal(s1[f"perm{l}"], s0[f"perm{l}"]), f"perm level {l}"
        assert torch.equal(s1[f"score{l}"], s0[f"score{l}"]), f"score level {l}"
        assert torch.equal(s1[f"h{l}"], s0[f"h{l}"]), f"h level {l}"
    assert torch.equal(e1, e0)
    # and the training forward (autograd path, kernel-per-operator) gives the same embeddings
    e2 = model(cb, None, c.node_ptr)
    assert torch.equal(e2.detach(), e0)


def test_fused_rejects_uncoalesced_lists(cuda):
    """An edge list that is not sorted / not symmetric must fail loudly (status word -> RuntimeError), never silently."""
    from tsg import nn as tnn
    c = synth.make_corpus("PROTEINS", 6, seed=3)
    c.row[[0, 1]] = c.row[[1, 0]]; c.col[[0, 1]] = c.col[[1, 0]]          # swap two edges of graph 0: unsorted
    cb = _compact(c, cuda)
    model = tnn.PackedSAGNet(c.num_node_labels, 32, 8, 0.5, 0.0).to(cuda)
    tnn.check_fused_status()
    with torch.no_grad():
        model(cb, None, c.node_ptr)
    with pytest.raises(RuntimeError, match="coalesced"):
        tnn.check_fused_status()


@pytest.mark.parametrize("shape,G", [("DD", 60), ("PROTEINS", 400), ("JANY", 10)])
def test_k1d_symmetric_csr_equals_k1b(cuda, shape, G):
    """K1d (one launch, one orientation, no atomics) == K1b's dst-major AND src-major CSR, bit for bit."""
    from tsg import _lib, ops
    from tsg._lib import call, ptr, stream_ptr
    c = synth.make_corpus(shape, G, seed=23)
    cb = _compact(c, cuda)
    N, E = int(c.node_ptr[-1]), int(c.edge_ptr[-1])
    nmax = int(np.diff(c.node_ptr).max())
    ref = ops.build_csr_graphs_local(cb, N, nmax)
    rowptr = torch.empty(N + 1, dtype=torch.int32, device=cuda)
    colidx = torch.empty(E + N, dtype=torch.int32, device=cuda)
    val = torch.empty(E + N, dtype=torch.float32, device=cuda)
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    call("tsg_csr_build_graphs_sym_local", ptr(cb.row), ptr(cb.col), ptr(cb.edge_ptr), ptr(cb.node_ptr), G, N, E, nmax,
         ptr(rowptr), ptr(colidx), ptr(val), ptr(status), stream_ptr())
    assert int(status.item()) == 0
    for mine, a, b in ((rowptr, ref.rowptr, ref.t_rowptr), (colidx, ref.colidx, ref.t_colidx), (val, ref.val, ref.t_val)):
        assert torch.equal(mine, a) and torch.equal(mine, b)


@pytest.mark.parametrize("shape,G,nhid", [("DD", 30, 32), ("PROTEINS", 200, 32), ("DD", 7, 128)])
def test_coalesced_step_is_bit_identical(cuda, shape, G, nhid):
    """CompactBatch.coalesced = True (K1d + single-orientation K1c, transposed CSR aliased) against the general kernels:
    same embeddings and same parameter gradients, bit for bit -- the arrays the kernels read are identical."""
    from tsg import nn as tnn
    c = synth.make_corpus(shape, G, seed=29)
    torch.manual_seed(2)
    model = tnn.PackedSAGNet(c.num_node_labels, nhid, 8, 0.5, 0.0).to(cuda)
    cot = torch.randn(G, 8, generator=torch.Generator().manual_seed(3)).to(cuda)
    res = []
    for flag in (False, True):
        cb = _compact(c, cuda)
        cb.coalesced = flag
        model.zero_grad(set_to_none=True)
        z = model(cb, None, c.node_ptr)
        (z * cot).sum().backward()
        tnn.check_fused_status()
        res.append((z.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0])
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k


def test_coalesced_promise_is_verified(cuda):
    from tsg import nn as tnn
    c = synth.make_corpus("PROTEINS", 6, seed=3)
    # drop one direction of an edge of graph 2: sorted, in range, but not symmetric
    e0 = int(c.edge_ptr[2])
    keep = np.ones(c.row.shape[0], bool); keep[e0] = False
    c.row, c.col = c.row[keep], c.col[keep]
    c.edge_ptr = c.edge_ptr.copy(); c.edge_ptr[3:] -= 1
    cb = _compact(c, cuda)
    cb.coalesced = True
    model = tnn.PackedSAGNet(c.num_node_labels, 32, 8, 0.5, 0.0).to(cuda)
    tnn.check_fused_status()
    model(cb, None, c.node_ptr)
    with pytest.raises(RuntimeError, match="symmetric"):
        tnn.check_fused_status()
