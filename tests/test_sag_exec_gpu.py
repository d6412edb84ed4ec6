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
