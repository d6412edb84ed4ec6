"""Launcher: run an UNMODIFIED reference script on top of tsg.

    python -m tsg.run /path/to/two-stage-gnn/Code/sag/train_triplet.py --dataset=DD --epochs=2 ...

None of it touches the reference files.  What it installs (every item is a defect or stack incompatibility listed in
SURVEY A.3, with the reference file:line it papers over):

  import surface
    * the `torch_geometric` shim (pyg_shim/) and the script's directory on sys.path;
    * `gen` package alias: Code/sage+gat+diffpool/train*.py:18-22 import gen.feat / gen.data but that directory ships
      no gen/ -- the sibling Code/eigengcn/gen is appended to the END of sys.path;
    * stub modules for `matplotlib`, `matplotlib.pyplot` and `community` (python-louvain) when they are not
      installed: eigengcn/graph.py:3 and coarsen_pooling_with_last_eigen_padding.py:3-5 import them and never use them;
    * `sklearn.neighbors.LSHForest` (eigengcn/graph.py:47, unused function) as a name that raises when called;
      `SpectralClustering.fit` accepts the np.matrix coarsen_pooling_with_last_eigen_padding.py:97,103 hands it;
    * `encoders_GAT.DGATHead_V3 = ()` via builtins (undefined upstream: every GAT construction raises as shipped,
      encoders_GAT.py:64-68).
  2020-era stack
    * networkx >= 3: `nx.to_numpy_matrix` (graph_sampler.py:28, cross_val.py:164,194,227), `Graph.node`
      (eigengcn/graph_sampler.py:52,79,99-100, eigengcn/train.py:335,339), and `float(nx.__version__)`
      (load_data.py:112 raises on '3.6.1': the version string is shortened to major.minor);
    * `os.environ[...] = <int>` (eigengcn/train.py:649 assigns the int default of --cuda): non-str values are str()ed;
    * `parser.set_defaults(pool_sizes=10)` for an option declared `type=str` (eigengcn/train*.py:551,614 -- the int
      default then meets `'...' + args.pool_sizes`, :262): defaults of `type=str` options are str()ed.
  device hygiene (the scripts mix CPU and CUDA tensors in their evaluation code)
    * `Tensor.numpy()` on a CUDA / grad tensor goes through `.detach().cpu()` (Code/sag/train_triplet.py:49,91);
    * `nn.Sequential.forward` moves a CPU input to the module's device, `F.cross_entropy` its target to the input's
      device and `Tensor.eq` its other operand to self's device (sag/train_triplet.py:104-130 feeds CPU tensors to a
      CUDA MLP and compares a CUDA prediction with a CPU label);
    * index assignment through a uint8 mask (an error since torch 1.2) is taken as the bool mask it meant
      (sage+gat+diffpool/encoders.py:436, the --linkpred loss).
  script text (compiled from an AST of the file, the file itself is only read)
    * `def evaluate(train_loader, val_loader, model, device)` called as `evaluate(a, b, model, name=..., max_num_examples=...)`
      (sag/train_triplet_pre_train.py:34 vs :242,277,282,285): the definition gains `device=None, **_ignored`, a missing
      device resolves to the model's.
  B2 patches
    * the script directory's `encoders` / `encoders_GAT` modules are pre-imported and their hot methods swapped for the
      tsg drop-ins (tsg/dense_patch.py).

Then the script runs with __name__ == "__main__" in the caller's working directory.
"""
from __future__ import annotations

import ast
import builtins
import importlib
import os
import sys
import types


# --------------------------------------------------------------------------------------------- import surface
def _stub_module(name: str, doc: str, **attrs):
    m = types.ModuleType(name, doc)
    m.__dict__.update(attrs)
    m.__tsg_stub__ = True

    def __getattr__(attr):           # any other attribute: a callable that explains itself when CALLED
        if attr.startswith("__"):
            raise AttributeError(attr)

        def missing(*a, **k):
            raise ModuleNotFoundError(f"{name}.{attr} was called, but {name} is not installed (tsg.run provides an "
                                      "import-only stub because the reference imports it without using it)")
        return missing
    m.__getattr__ = __getattr__
    sys.modules[name] = m
    return m


def _install_stubs() -> list:
    done = []
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub_module("matplotlib", "import-only stub (tsg.run)")
        mpl.pyplot = _stub_module("matplotlib.pyplot", "import-only stub (tsg.run)")
        mpl.colors = _stub_module("matplotlib.colors", "import-only stub (tsg.run)")
        mpl.use = lambda *a, **k: None
        done.append("matplotlib")
    try:
        import community  # noqa: F401
    except Exception:
        _stub_module("community", "import-only stub for python-louvain (tsg.run)")
        done.append("community")
    try:
        import sklearn.neighbors as skn
        if not hasattr(skn, "LSHForest"):
            class LSHForest:                                  # eigengcn/graph.py:47 (unused function upstream)
                def __init__(self, *a, **k):
                    raise NotImplementedError("sklearn.neighbors.LSHForest was removed in scikit-learn 0.21")
            skn.LSHForest = LSHForest
            done.append("sklearn.neighbors.LSHForest")
    except Exception:
        pass
    try:      # scikit-learn >= 1.2 rejects np.matrix; coarsen_pooling_with_last_eigen_padding.py:97,103 fits one
        import numpy as np
        from sklearn.cluster import SpectralClustering
        if not getattr(SpectralClustering.fit, "_tsg", False):
            _fit = SpectralClustering.fit

            def fit(self, X, y=None):
                return _fit(self, np.asarray(X) if isinstance(X, np.matrix) else X, y)
            fit._tsg = True
            SpectralClustering.fit = fit
            done.append("SpectralClustering.fit(np.matrix)")
    except Exception:
        pass
    return done


def _install_gen_alias(script_dir: str | None) -> bool:
    """sage+gat+diffpool imports `gen.feat` / `gen.data` but only Code/eigengcn ships gen/."""
    if not script_dir or os.path.isdir(os.path.join(script_dir, "gen")):
        return False
    sibling = os.path.join(os.path.dirname(script_dir), "eigengcn")
    if os.path.isdir(os.path.join(sibling, "gen")):
        if sibling not in sys.path:
            sys.path.append(sibling)             # at the END: only `gen` resolves there, nothing shadows the script dir
        return True
    return False


# --------------------------------------------------------------------------------------------- stack compat
def _install_networkx_compat() -> None:
    try:
        import networkx as nx
        import numpy as np
    except Exception:
        return
    if not hasattr(nx, "to_numpy_matrix"):
        nx.to_numpy_matrix = lambda G, *a, **k: np.asmatrix(nx.to_numpy_array(G, *a, **k))
    if not hasattr(nx.Graph, "node"):
        nx.Graph.node = property(lambda self: self.nodes)
    try:
        float(nx.__version__)
    except ValueError:                                        # load_data.py:112: float(nx.__version__)
        nx.__version__ = ".".join(nx.__version__.split(".")[:2])


def _install_environ_compat() -> None:
    cls = type(os.environ)
    if getattr(cls.__setitem__, "_tsg", False):
        return
    orig = cls.__setitem__

    def __setitem__(self, key, value):                        # eigengcn/train.py:649: os.environ[...] = args.cuda (int)
        return orig(self, key, value if isinstance(value, (str, bytes)) else str(value))
    __setitem__._tsg = True
    cls.__setitem__ = __setitem__


def _install_argparse_compat() -> None:
    import argparse
    if getattr(argparse.ArgumentParser.set_defaults, "_tsg", False):
        return
    orig = argparse.ArgumentParser.set_defaults

    def set_defaults(self, **kwargs):
        for a in self._actions:
            if a.dest in kwargs and a.type is str and kwargs[a.dest] is not None and not isinstance(kwargs[a.dest], str):
                kwargs[a.dest] = str(kwargs[a.dest])
        return orig(self, **kwargs)
    set_defaults._tsg = True
    argparse.ArgumentParser.set_defaults = set_defaults


def _install_device_hygiene() -> None:
    import torch
    import torch.nn.functional as F
    if not getattr(torch.Tensor.numpy, "_tsg", False):
        _numpy = torch.Tensor.numpy

        def numpy(self, *a, **k):
            return _numpy(self.detach().cpu() if self.is_cuda or self.requires_grad else self, *a, **k)
        numpy._tsg = True
        torch.Tensor.numpy = numpy
    if not getattr(torch.nn.Sequential.forward, "_tsg", False):
        _seq = torch.nn.Sequential.forward

        def forward(self, input):
            if torch.is_tensor(input) and not input.is_cuda:
                p = next(self.parameters(), None)
                if p is not None and p.is_cuda:
                    input = input.to(p.device)
            return _seq(self, input)
        forward._tsg = True
        torch.nn.Sequential.forward = forward
    if not getattr(F.cross_entropy, "_tsg", False):
        _ce = F.cross_entropy

        def cross_entropy(input, target, *a, **k):
            if torch.is_tensor(target) and target.device != input.device:
                target = target.to(input.device)
            return _ce(input, target, *a, **k)
        cross_entropy._tsg = True
        F.cross_entropy = cross_entropy
    if not getattr(torch.Tensor.__setitem__, "_tsg", False):
        _setitem = torch.Tensor.__setitem__

        def __setitem__(self, idx, val):             # encoders.py:436: link_loss[1 - adj_mask.byte()] = 0.0 (uint8 mask)
            if getattr(idx, "dtype", None) == torch.uint8:
                idx = idx.bool()
            return _setitem(self, idx, val)
        __setitem__._tsg = True
        torch.Tensor.__setitem__ = __setitem__
    if not getattr(torch.Tensor.eq, "_tsg", False):
        _eq = torch.Tensor.eq

        def eq(self, other):
            if torch.is_tensor(other) and other.device != self.device:
                other = other.to(self.device)
            return _eq(self, other)
        eq._tsg = True
        torch.Tensor.eq = eq


def install(script_dir: str | None = None) -> dict:
    """Everything except the B2 patches and the script itself.  Returns what was installed (for tests / logging)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(here, "pyg_shim")
    for p in (shim, here):
        if p not in sys.path:
            sys.path.insert(0, p)
    if script_dir and script_dir not in sys.path:
        sys.path.insert(0, script_dir)
    report = {"stubs": _install_stubs(), "gen_alias": _install_gen_alias(script_dir)}
    _install_device_hygiene()
    _install_networkx_compat()
    _install_environ_compat()
    _install_argparse_compat()
    if not hasattr(builtins, "DGATHead_V3"):
        builtins.DGATHead_V3 = ()
    return report


# --------------------------------------------------------------------------------------------- B2 patches
def patch_dense_modules(script_dir: str):
    """B2: pre-import the script directory's `encoders` / `encoders_GAT` modules (they then sit in
    sys.modules for the script's own imports) and swap the hot methods for the tsg drop-ins."""
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


# --------------------------------------------------------------------------------------------- script text
class _FixEvaluateSignature(ast.NodeTransformer):
    """`def evaluate(train_loader, val_loader, model, device)` -> `(..., device=None, **_tsg_ignored)` with
    `device` resolved from the model when missing (sag/train_triplet_pre_train.py:34 vs its call sites)."""

    def __init__(self):
        self.fixed = []

    def visit_FunctionDef(self, node: ast.FunctionDef):
        self.generic_visit(node)
        names = [a.arg for a in node.args.args]
        if (node.name == "evaluate" and names == ["train_loader", "val_loader", "model", "device"]
                and not node.args.defaults and node.args.kwarg is None):
            node.args.defaults = [ast.Constant(None)]
            node.args.kwarg = ast.arg("_tsg_ignored")
            fix = ast.parse("if device is None:\n    device = next(model.parameters()).device").body[0]
            node.body.insert(0, fix)
            self.fixed.append(node.lineno)
        return node


def compile_script(path: str):
    """(code object, list of fixes applied).  The file is read, never written."""
    with open(path, "rb") as f:
        src = f.read()
    tree = ast.parse(src, filename=path)
    fx = _FixEvaluateSignature()
    tree = fx.visit(tree)
    ast.fix_missing_locations(tree)
    return compile(tree, path, "exec"), [f"evaluate() signature at line {ln}" for ln in fx.fixed]


def run_script(script: str, argv: list) -> dict:
    """Execute the (compat-compiled) script as __main__; returns its globals."""
    code, fixes = compile_script(script)
    mod = types.ModuleType("__main__")
    mod.__file__ = script
    mod.__tsg_fixes__ = fixes
    old_main, old_argv = sys.modules.get("__main__"), sys.argv
    sys.modules["__main__"] = mod
    sys.argv = [script] + list(argv)
    try:
        exec(code, mod.__dict__)
    finally:
        sys.argv = old_argv
        if old_main is not None:
            sys.modules["__main__"] = old_main
    return mod.__dict__


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m tsg.run <reference script.py> [script args...]")
    script = os.path.abspath(argv[0])
    install(os.path.dirname(script))
    patch_dense_modules(os.path.dirname(script))
    # the working directory stays the caller's: the scripts' relative paths (`data/<NAME>`, `latest.pth`) resolve
    # exactly as under `python train.py` run from that directory
    run_script(script, argv[1:])


if __name__ == "__main__":
    main()
