import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "two-stage-gnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def rel_err(a, b):
    """L_inf(a-b) / max(L_inf(b), tiny): the per-tensor relative error of SURVEY.md section 4."""
    import torch
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    if a.numel() == 0:
        return 0.0
    e = float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
    if os.environ.get("TSG_ERRLOG"):
        _errlog("rel", e, str(tuple(a.shape)))
    return e


def _errlog(kind, value, extra=""):
    """TSG_ERRLOG=<file>: append every measured relative error with the test that measured it (how the tolerance table
    in profiles/ is produced on the GPU box)."""
    path = os.environ.get("TSG_ERRLOG")
    if path:
        with open(path, "a") as f:
            f.write(f"{os.environ.get('PYTEST_CURRENT_TEST', '?')}\t{kind}\t{value:.3e}\t{extra}\n")


def grad_check(got, ref32, ref64=None, tol=1e-5, name=""):
    """north_star gate for a gradient: within `tol` (1e-5 relative) of the fp32 reference.  Where the quantity is
    ill-conditioned in the REFERENCE's own arithmetic, the reference's fp32 result itself sits further than `tol`
    from its float64 evaluation; then (and only when a float64 reference is supplied) the gate is "at least as close
    to float64 as 3x the fp32 reference is", and the conditioning number is part of the assertion message."""
    e = rel_err(got, ref32)
    _errlog("grad", e, name)
    if e <= tol:
        return e
    if ref64 is not None:
        cond = rel_err(ref32, ref64)
        e64 = rel_err(got, ref64)
        _errlog("grad64", e64, f"{name} cond={cond:.3e}")
        assert e64 <= max(tol, 3.0 * cond), (f"{name}: |cuda - ref32| = {e:.2e}, |cuda - ref64| = {e64:.2e}, "
                                             f"reference fp32-vs-fp64 conditioning = {cond:.2e}")
        return e64
    raise AssertionError(f"{name}: gradient off by {e:.2e} (> {tol:.0e}) and no float64 conditioning reference")
