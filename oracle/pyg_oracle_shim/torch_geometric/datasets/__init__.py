"""TUDataset: the product shim's reader (native TU loader, host code) executed under this package."""
import importlib.util
import os
import sys

_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), *[".."] * 4, "two-stage-gnn_b200", "pyg_shim",
                    "torch_geometric", "datasets", "__init__.py")
_spec = importlib.util.spec_from_file_location("torch_geometric._datasets_impl", os.path.normpath(_SRC))
_impl = importlib.util.module_from_spec(_spec)
sys.modules["torch_geometric._datasets_impl"] = _impl
_spec.loader.exec_module(_impl)
TUDataset = _impl.TUDataset
