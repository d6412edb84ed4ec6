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


def grad_check(got, ref32, ref64=None, tol=1e-5, name="", fwd_noise=1e-6):
    """north_star gate for a gradient: within `tol` (1e-5 relative) of the fp32 reference.  A few quantities are
    ill-conditioned in the REFERENCE's own arithmetic (node-BN rows whose variance is comparable with eps = 1e-5
    amplify rounding noise by 1/sqrt(var + eps); the hinge gradient at random init is a difference of near-equal unit
    vectors): there the reference's fp32 result itself sits far beyond `tol` from its float64 evaluation.  For those --
    and only when a float64 reference is supplied -- the gate scales with the MEASURED amplification: with
    cond = |ref32 - ref64| (the reference's own fp32 noise on this quantity) and fwd_noise = the reference's fp32
    noise on a well-conditioned quantity (its forward output, ~1e-6), the quantity amplifies noise by cond / fwd_noise,
    and the CUDA result may sit tol * cond / fwd_noise from float64 (i.e. 1e-5 before amplification).  Both numbers
    are part of the assertion message and of the TSG_ERRLOG table."""
    e = rel_err(got, ref32)
    _errlog("grad", e, name)
    if e <= tol:
        return e
    if ref64 is not None:
        cond = rel_err(ref32, ref64)
        e64 = rel_err(got, ref64)
        allowed = max(tol, tol * cond / max(fwd_noise, 1e-7))
        _errlog("grad64", e64, f"{name} cond={cond:.3e} allowed={allowed:.3e}")
        assert e64 <= allowed, (f"{name}: |cuda - ref32| = {e:.2e}, |cuda - ref64| = {e64:.2e}, reference fp32-vs-fp64 "
                                f"conditioning = {cond:.2e}, forward fp32 noise = {fwd_noise:.2e}, allowed = {allowed:.2e}")
        return e64
    raise AssertionError(f"{name}: gradient off by {e:.2e} (> {tol:.0e}) and no float64 conditioning reference")
