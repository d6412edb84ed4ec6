"""torch_geometric.nn surface used by Code/sag/network.py:2-4 and layers.py:1."""
import torch

from tsg import ops
from tsg.nn import GCNConv  # noqa: F401
from . import pool  # noqa: F401


def _graph_ptr(batch: torch.Tensor, size):
    # PyG: size = batch.max().item() + 1 (one host sync, as upstream)
    G = int(batch.max().item()) + 1 if size is None else int(size)
    return ops.batch_to_ptr(batch, G)


def global_max_pool(x, batch, size=None):
    """network.py:36,40,44 `gmp`."""
    return ops.readout(x, _graph_ptr(batch, size), ops.READOUT_MAX)


def global_mean_pool(x, batch, size=None):
    """network.py:36,40,44 `gap`."""
    return ops.readout(x, _graph_ptr(batch, size), ops.READOUT_MEAN)


def global_add_pool(x, batch, size=None):
    return ops.readout(x, _graph_ptr(batch, size), ops.READOUT_SUM)


class _ImportableOnly(torch.nn.Module):
    """network.py:3 imports GraphConv and TopKPooling but never instantiates them."""

    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__} is importable for compatibility only; "
                                  "the reference never constructs it")


class GraphConv(_ImportableOnly):
    pass


class TopKPooling(_ImportableOnly):
    pass
