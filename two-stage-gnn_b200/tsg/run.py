"""Launcher: run an UNMODIFIED reference script on top of tsg.

    python -m tsg.run /path/to/two-stage-gnn/Code/sag/train_triplet.py --dataset=DD --epochs=2 ...

What it does (none of it touches the reference files):
  * puts the `torch_geometric` shim (pyg_shim/) and the script's directory on sys.path;
  * compat aliases for the 2020-era stack the scripts were written for: `Tensor.numpy()` on a CUDA
    tensor goes through `.cpu()` (Code/sag/train_triplet.py:49,91 call it on device tensors);
    `encoders_GAT.DGATHead_V3 = ()` (undefined upstream, every GAT construction raises as shipped);
    `nx.to_numpy_matrix` / `Graph.node` for networkx >= 3;
  * runpy.run_path(script, run_name="__main__").
"""
from __future__ import annotations

import builtins
import os
import runpy
import sys


def install(script_dir: str | None = None) -> None:
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(here, "pyg_shim")
    for p in (shim, here):
        if p not in sys.path:
            sys.path.insert(0, p)
    if script_dir and script_dir not in sys.path:
        sys.path.insert(0, script_dir)
    import torch
    if not getattr(torch.Tensor.numpy, "_tsg", False):
        _numpy = torch.Tensor.numpy

        def numpy(self, *a, **k):
            return _numpy(self.detach().cpu() if self.is_cuda or self.requires_grad else self, *a, **k)
        numpy._tsg = True
        torch.Tensor.numpy = numpy
    if not hasattr(builtins, "DGATHead_V3"):
        builtins.DGATHead_V3 = ()
    try:
        import networkx as nx
        import numpy as np
        if not hasattr(nx, "to_numpy_matrix"):
            nx.to_numpy_matrix = lambda G, *a, **k: np.asmatrix(nx.to_numpy_array(G, *a, **k))
        if not hasattr(nx.Graph, "node"):
            nx.Graph.node = property(lambda self: self.nodes)
    except Exception:
        pass


def patch_dense_modules(script_dir: str):
    """B2: pre-import the script directory's `encoders` / `encoders_GAT` modules (they then sit in
    sys.modules for the script's own imports) and swap the hot methods for the tsg drop-ins."""
    import importlib
    from . import dense_patch
    done = {}
    if os.path.exists(os.path.join(script_dir, "encoders.py")):
        enc = importlib.import_module("encoders")
        if hasattr(enc, "Pool"):
            done.update(dense_patch.install(eigen_encoders=enc))
        else:
            done.update(dense_patch.install(encoders=enc))
    if os.path.exists(os.path.join(script_dir, "encoders_GAT.py")):
        done.update(dense_patch.install(encoders_gat=importlib.import_module("encoders_GAT")))
    return done


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m tsg.run <reference script.py> [script args...]")
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    patch_dense_modules(os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    # the working directory stays the caller's: the scripts' relative paths (`data/<NAME>`, `latest.pth`) resolve
    # exactly as under `python train.py` run from that directory
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
