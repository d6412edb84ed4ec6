"""CUDA-graph capture of a whole training step through the C ABI (tsg.train.CapturedStep): replays reproduce the eager
step bit for bit -- loss sequence and final parameters -- for the dense GraphSAGE/base encoder (K2 RAW + K3 fused
epilogue + K6 + the last-CTA reduction tickets + memsets inside the library all captured)."""
import numpy as np
import pytest
import torch

from tsg import synth

pytestmark = pytest.mark.gpu


def _setup(cuda, seed):
    from tsg import dense, ops
    corpus = synth.make_corpus("PROTEINS", 64, seed=11)
    pk = synth.pack(corpus, one_hot=False)
    n = np.diff(corpus.node_ptr)
    g = torch.Generator().manual_seed(1)
    table = torch.randn(700, 16, generator=g)
    x = table[torch.from_numpy(np.concatenate([np.arange(k) for k in n]))].to(cuda)
    ei = torch.from_numpy(pk["edge_index"]).to(cuda)
    N = int(corpus.node_ptr[-1])
    csr = ops.build_csr(ops.EdgeList.from_edge_index(ei), N, mode=ops.CSR_RAW)
    gptr = torch.from_numpy(corpus.node_ptr).to(cuda)
    has_pad = torch.ones(corpus.num_graphs, dtype=torch.bool, device=cuda)
    y = torch.from_numpy(corpus.y).to(cuda)
    torch.manual_seed(seed)
    model = dense.PackedGcnEncoder(16, 16, 16, 2, 3, bn=True, final_dim="number_classes").to(cuda)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True)

    def body():
        _, logits = model(x, csr, gptr, has_pad)
        loss = torch.nn.functional.cross_entropy(logits, y)
        opt.zero_grad(set_to_none=False); loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 2.0)
        opt.step()
        return loss.detach()
    return model, body


def test_captured_step_equals_eager(cuda):
    from tsg.train import CapturedStep
    steps, warm = 6, 3
    model_e, body_e = _setup(cuda, 5)
    eager = [float(body_e()) for _ in range(warm + 1 + steps)]          # warm-up + the captured call + replays
    model_c, body_c = _setup(cuda, 5)
    cap = CapturedStep(body_c, warmup=warm)                              # runs body warm times eagerly, once capturing (not executed)
    captured = [float(cap.replay()) for _ in range(steps + 1)]
    assert captured == eager[warm:warm + steps + 1]
    for (k, a), (_, b) in zip(model_e.state_dict().items(), model_c.state_dict().items()):
        assert torch.equal(a, b), k
    assert captured[-1] < captured[0]
