"""Data / Batch / DataLoader: host-side containers, no operators -- the product shim's own file is executed here
(pyg_shim/torch_geometric/data/__init__.py), so scripts run end to end on CPU over the oracle operators."""
import importlib.util
import os

_SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), *[".."] * 4, "two-stage-gnn_b200", "pyg_shim",
                    "torch_geometric", "data", "__init__.py")
_spec = importlib.util.spec_from_file_location(__name__ + "._impl", os.path.normpath(_SRC))
_impl = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_impl)
Data, Batch, DataLoader = _impl.Data, _impl.Batch, _impl.DataLoader
