"""Drop-in `torch_geometric` package exposing exactly the PyG 1.x names that the reference's
Code/sag imports (layers.py:1-2, network.py:2-4, train*.py:5-7), backed by libtsg.so.

Put `two-stage-gnn_b200/pyg_shim` on sys.path (the launcher `python -m tsg.run` does) and the
reference scripts import this instead of PyTorch-Geometric.  No CPU fallback: CPU tensors raise.
"""
__version__ = "1.6.3+tsg"
from . import data, datasets, nn, utils  # noqa: F401,E402
