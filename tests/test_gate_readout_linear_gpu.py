"""tsg_gate_readout_linear_fwd (SAGPool gate + the next GCNConv's x W in one flat pass, then the [gmp || gap] readout; K6)
against the three-kernel sequence it replaces -- tsg_gate_gather_fwd, tsg_readout_fwd, tsg_linear_fwd -- BIT FOR BIT (the
header promises it), on ragged graphs around the kernels' batch sizes, and against the oracle (Code/sag/layers.py:21,
network.py:36) within the float tolerance."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5
READOUT_MAX, READOUT_MEAN = 1, 2


def _case(sizes_in, ratio, F, seed, dev):
    """graphs of sizes_in nodes; perm = a random k-subset per graph in a random order (what top-k hands over)"""
    rng = np.random.default_rng(seed)
    n_in = np.asarray(sizes_in, np.int64)
    k = np.where(n_in > 0, np.ceil(ratio * n_in).astype(np.int64), 0)
    ptr_in = np.concatenate([[0], np.cumsum(n_in)]); ptr_out = np.concatenate([[0], np.cumsum(k)])
    perm = np.concatenate([ptr_in[g] + rng.permutation(n_in[g])[:k[g]] for g in range(len(n_in))] + [np.zeros(0, np.int64)])
    N = int(ptr_in[-1])
    x = rng.standard_normal((N, F)).astype(np.float32)
    x[rng.random((N, F)) < 0.2] = 0.0                              # ReLU output: exact zeros, ties in the max
    score = rng.standard_normal(N).astype(np.float32)
    W = (rng.standard_normal((F, F)) / np.sqrt(F)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t(x), t(score), t(perm.astype(np.int64)), t(ptr_out.astype(np.int64)), t(W), ptr_out


def _sequence(x, score, perm, gptr, W, G, K, F):
    from tsg import _lib
    xo = torch.empty(K, F, device=x.device); out = torch.empty(G, 2 * F, device=x.device)
    am = torch.empty(G, F, dtype=torch.int32, device=x.device); xw = torch.empty(K, F, device=x.device)
    if K:
        _lib.call("tsg_gate_gather_fwd", _lib.ptr(x), _lib.ptr(score), _lib.ptr(perm), None, _lib.ptr(xo), None, K, F, _lib.stream_ptr())
    _lib.call("tsg_readout_fwd", _lib.ptr(xo), _lib.ptr(gptr), G, F, READOUT_MAX | READOUT_MEAN, _lib.ptr(out), 2 * F, _lib.ptr(am), _lib.stream_ptr())
    if K:
        _lib.call("tsg_linear_fwd", _lib.ptr(xo), _lib.ptr(W), None, _lib.ptr(xw), K, F, F, 0, 0, _lib.stream_ptr())
    return xo, out, am, xw


def _fused(x, score, perm, gptr, W, G, K, F, linear=True):
    from tsg import _lib
    xo = torch.full((K, F), float("nan"), device=x.device); out = torch.full((G, 2 * F), float("nan"), device=x.device)
    am = torch.full((G, F), -7, dtype=torch.int32, device=x.device); xw = torch.full((K, F), float("nan"), device=x.device)
    _lib.call("tsg_gate_readout_linear_fwd", _lib.ptr(x), _lib.ptr(score), _lib.ptr(perm), _lib.ptr(gptr), G, K, F, _lib.ptr(xo),
              _lib.ptr(out), 2 * F, _lib.ptr(am), _lib.ptr(W) if linear else None, _lib.ptr(xw) if linear else None, _lib.stream_ptr())
    return xo, out, am, xw


SIZES = [1, 2, 7, 0, 126, 127, 128, 129, 130, 255, 256, 257, 5, 1000, 64, 3, 5748, 0, 31, 383]


@pytest.mark.parametrize("F", [32, 64, 16, 128])        # 32 / 64: k_gate_linear; 16 / 128: the documented fallback
@pytest.mark.parametrize("linear", [True, False])
def test_bit_identical_to_the_three_kernel_sequence(cuda, F, linear):
    x, score, perm, gptr, W, ptr_out = _case(SIZES, 0.5, F, 3, cuda)
    G, K = len(SIZES), int(ptr_out[-1])
    ref = _sequence(x, score, perm, gptr, W, G, K, F)
    got = _fused(x, score, perm, gptr, W, G, K, F, linear)
    names = ("xo", "out", "argmax", "xw")
    for name, r, g_ in zip(names, ref, got):
        if name == "xw" and not linear:
            assert torch.isnan(g_).all(), "xw_next must not be touched without w_next"
            continue
        assert torch.equal(r, g_), f"{name} differs (F={F})"
        if r.dtype == torch.float32:        # torch.equal treats -0 == +0; the header says bit-identical
            assert torch.equal(r.view(torch.int32), g_.view(torch.int32)), f"{name}: sign of zero differs (F={F})"


def test_many_small_graphs_and_grid_stride(cuda):
    """more graphs than warps in the readout grid: every warp walks several graphs"""
    rng = np.random.default_rng(5)
    sizes = rng.integers(0, 40, 6000).tolist()
    x, score, perm, gptr, W, ptr_out = _case(sizes, 0.5, 32, 9, cuda)
    G, K = len(sizes), int(ptr_out[-1])
    ref = _sequence(x, score, perm, gptr, W, G, K, 32)
    got = _fused(x, score, perm, gptr, W, G, K, 32)
    for r, g_ in zip(ref, got):
        assert torch.equal(r, g_)


def test_against_the_oracle(cuda):
    """layers.py:21 x[perm] * tanh(score[perm]), network.py:36 [gmp || gap], GCNConv's x @ W"""
    from oracle import pyg_ref as R
    sizes = [40, 1, 300, 77]
    x, score, perm, gptr, W, ptr_out = _case(sizes, 0.5, 32, 11, cuda)
    G, K = len(sizes), int(ptr_out[-1])
    xo, out, am, xw = _fused(x, score, perm, gptr, W, G, K, 32)
    xc, sc, pc = x.cpu(), score.cpu(), perm.cpu()
    xo_ref = xc[pc] * torch.tanh(sc[pc]).view(-1, 1)
    batch = torch.repeat_interleave(torch.arange(G), torch.from_numpy(np.diff(ptr_out)))
    ro_ref = torch.cat([R.global_max_pool(xo_ref, batch, G), R.global_mean_pool(xo_ref, batch, G)], 1)
    assert rel_err(xo, xo_ref) <= TOL and rel_err(out, ro_ref) <= TOL and rel_err(xw, xo_ref @ W.cpu()) <= TOL
    # the argmax rows hold the max (first one on ties)
    for g in range(G):
        rows = am[g].cpu().long()
        assert torch.equal(xo.cpu()[rows, torch.arange(32)], out[g, :32].cpu())


def test_argument_checks(cuda):
    from tsg import _lib
    x, score, perm, gptr, W, ptr_out = _case([4, 4], 0.5, 32, 1, cuda)
    xo = torch.empty(4, 32, device=cuda); out = torch.empty(2, 64, device=cuda); am = torch.empty(2, 32, dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="go together"):
        _lib.call("tsg_gate_readout_linear_fwd", _lib.ptr(x), _lib.ptr(score), _lib.ptr(perm), _lib.ptr(gptr), 2, 4, 32, _lib.ptr(xo),
                  _lib.ptr(out), 64, _lib.ptr(am), _lib.ptr(W), None, _lib.stream_ptr())
    with pytest.raises(RuntimeError, match="bad shape"):
        _lib.call("tsg_gate_readout_linear_fwd", _lib.ptr(x), _lib.ptr(score), _lib.ptr(perm), _lib.ptr(gptr), 2, 4, 30, _lib.ptr(xo),
                  _lib.ptr(out), 64, _lib.ptr(am), None, None, _lib.stream_ptr())
