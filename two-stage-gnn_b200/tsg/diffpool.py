"""DiffPool (Code/sage+gat+diffpool/encoders.py:236-406, SoftPoolingGcnEncoder, num_pooling = 1) on
the packed layout.  State-dict keys match the reference (`assign_conv_first_modules.0.weight`,
`assign_pred_modules.0.bias`, `conv_first_after_pool.0.weight`, ...).

    Z  = gcn_forward(x, A)            (K2 + K3, masked rows do not exist in the packed layout)
    S  = softmax(Linear(gcn_forward_assign(x, A)))          [sum n, K],  K = int(max_nodes * ratio)
    T  = A S                          (K2 SpMM, F = K)
    X' = S^T Z, A' = S^T T            (K7: per-graph contractions on tcgen05, 3xTF32)
    Z' = gcn_forward(X', A')          (dense K x K weighted adjacency: K7 seg_linear + K3)
    out = [max_n Z (incl. zero padded rows) | max_K Z'] -> map_model
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .dense import GcnStack, _pred_layers, _readout_max, gcn_forward
from .ops import CSR, LIN_NODEBN, LIN_RELU, READOUT_MAX


import os
SPLIT_CONTRACT = os.environ.get("TSG_DIFFPOOL_ONE_CONTRACT", "0") != "1"


def dense_gcn_forward(x, adj, graph_ptr, convs, bn=True):
    """gcn_forward (encoders.py:140-167) when the adjacency is a per-graph DENSE K x K block
    (`adj` [G, K, K], x [G*K, F]): adj @ x is a per-graph row-local product (K7 seg_linear)."""
    G, K, _ = adj.shape
    rows = adj.reshape(G * K, K)
    outs = []
    for i, c in enumerate(convs):
        last = i == len(convs) - 1
        y = ops.seg_linear(rows, x.view(G, K, -1), graph_ptr)                # adj @ x
        flags = ops.LIN_NORMALIZE | (0 if last else (LIN_RELU | (LIN_NODEBN if bn else 0)))
        x = ops.linear(y, c.weight, c.bias, flags)
        outs.append(x)
    return torch.cat(outs, dim=1)


class PackedSoftPoolEncoder(nn.Module):
    def __init__(self, max_num_nodes, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 assign_hidden_dim, assign_ratio=0.25, pred_hidden_dims=(), bn=True, final_dim="output_dim"):
        super().__init__()
        self.bn, self.final_dim, self.max_num_nodes = bn, final_dim, max_num_nodes
        s = GcnStack(input_dim, hidden_dim, embedding_dim, num_layers)
        self.conv_first, self.conv_block, self.conv_last = s.conv_first, s.conv_block, s.conv_last
        D = hidden_dim * (num_layers - 1) + embedding_dim
        self.pred_input_dim = D
        s2 = GcnStack(D, hidden_dim, embedding_dim, num_layers)
        self.conv_first_after_pool = nn.ModuleList([s2.conv_first])
        self.conv_block_after_pool = nn.ModuleList([s2.conv_block])
        self.conv_last_after_pool = nn.ModuleList([s2.conv_last])
        self.assign_dim = int(max_num_nodes * assign_ratio)                  # encoders.py:283
        s3 = GcnStack(input_dim, assign_hidden_dim, self.assign_dim, num_layers)
        self.assign_conv_first_modules = nn.ModuleList([s3.conv_first])
        self.assign_conv_block_modules = nn.ModuleList([s3.conv_block])
        self.assign_conv_last_modules = nn.ModuleList([s3.conv_last])
        apd = assign_hidden_dim * (num_layers - 1) + self.assign_dim
        self.assign_pred_modules = nn.ModuleList([nn.Linear(apd, self.assign_dim)])
        self.pre_pred_model = _pred_layers(D * 2, pred_hidden_dims, embedding_dim)
        self.pred_model = _pred_layers(embedding_dim, pred_hidden_dims, label_dim)
        self.map_model = _pred_layers(D * 2, pred_hidden_dims, embedding_dim)
        self.map2_model = _pred_layers(embedding_dim, (), 2)

    @staticmethod
    def _stack(first, block, last):
        return [first] + list(block) + [last]

    def readout(self, x, csr: CSR, graph_ptr, has_pad, return_aux=False):
        G, K = graph_ptr.numel() - 1, self.assign_dim
        convs = self._stack(self.conv_first, self.conv_block, self.conv_last)
        z = gcn_forward(x, csr, convs, self.bn)                                              # :350
        out0 = _readout_max(z, graph_ptr, 0.0, has_pad)                                      # :353
        aconvs = self._stack(self.assign_conv_first_modules[0], self.assign_conv_block_modules[0],
                             self.assign_conv_last_modules[0])
        za = gcn_forward(x, csr, aconvs, self.bn)                                            # :365
        ap = self.assign_pred_modules[0]
        s = ops.linear(za, ap.weight.t(), ap.bias, ops.LIN_SOFTMAX)                          # :366-369: Linear + softmax in K3's epilogue
        t = ops.spmm(csr, s)                                                                 # adj @ S
        D = z.size(1)
        if SPLIT_CONTRACT:
            # two contractions, no copies: the single S^T [Z | AS] product needs torch.cat before it, two strided slices +
            # .contiguous() after it and their zero-fill + copy backward -- ~1.4 ms of ATen copies per step at config-4
            # size against ~0.2 ms for splitting S into hi / lo TF32 parts twice
            xp = ops.seg_contract(s, z, graph_ptr)                                           # :374  x = S^T Z
            apool = ops.seg_contract(s, t, graph_ptr)                                        # :375  adj = S^T (A S)
        else:
            c = ops.seg_contract(s, torch.cat([z, t], dim=1), graph_ptr)                     # :374-375 in one product
            xp, apool = c[:, :, :D].contiguous(), c[:, :, D:].contiguous()
        gptr2 = torch.arange(G + 1, device=x.device, dtype=torch.int64) * K
        convs2 = self._stack(self.conv_first_after_pool[0], self.conv_block_after_pool[0],
                             self.conv_last_after_pool[0])
        z2 = dense_gcn_forward(xp.view(G * K, D), apool, gptr2, convs2, self.bn)             # :378
        out1 = ops.readout(z2, gptr2, READOUT_MAX)                                           # :383
        out = torch.cat([out0, out1], dim=1)
        return (out, dict(s=s, xp=xp, ap=apool)) if return_aux else out

    def link_loss(self, s, csr: CSR, graph_ptr, num_nodes_host):
        """The `--linkpred` term of SoftPoolingGcnEncoder.loss (encoders.py:416-440) for the assignment `s` returned by
        readout(..., return_aux=True)["s"]; `num_nodes_host` = graph sizes (numpy / list) for the normaliser sum n^2."""
        import numpy as np
        n = np.asarray(num_nodes_host, dtype=np.float64)
        return ops.linkpred_loss(s, graph_ptr, csr, float((n * n).sum()))

    def forward(self, x, csr: CSR, graph_ptr, has_pad):
        output = self.readout(x, csr, graph_ptr, has_pad)
        if self.final_dim == "pretrain":
            out = self.map_model(output)
            return self.map2_model(out), out
        if self.final_dim != "output_dim":
            ov = self.pre_pred_model(output)
            return ov, self.pred_model(ov)
        return output, self.map_model(output)
