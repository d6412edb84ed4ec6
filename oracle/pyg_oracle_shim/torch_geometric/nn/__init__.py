"""Names Code/sag/network.py:2-4 and layers.py:1 import, backed by oracle.pyg_ref."""
import math

import torch

from oracle import pyg_ref as R
from . import pool  # noqa: F401


class GCNConv(torch.nn.Module):
    """PyG 1.6.3 GCNConv(in, out) defaults: weight [in, out] glorot-uniform, bias [out] zeros."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        a = math.sqrt(6.0 / (in_channels + out_channels))
        self.weight = torch.nn.Parameter((torch.rand(in_channels, out_channels) * 2 - 1) * a)
        self.bias = torch.nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index, edge_weight=None):
        return R.gcn_conv(x, edge_index, self.weight, self.bias, edge_weight)


def global_max_pool(x, batch, size=None):
    return R.global_max_pool(x, batch, size)


def global_mean_pool(x, batch, size=None):
    return R.global_mean_pool(x, batch, size)


class GraphConv(torch.nn.Module):        # network.py:3 imports the name, never constructs it
    pass


class TopKPooling(torch.nn.Module):
    pass
