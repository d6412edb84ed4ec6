"""Dense-GAT (Code/sage+gat+diffpool/encoders_GAT.py) on the packed CSR layout.

`PackedGatEncoder` mirrors DGATEncoderGraph (encoders_GAT.py:86-198): L DGATLayers of several
DGATHeads, hidden layers concatenate ELU(head) outputs, the last layer averages the heads then ELU,
max-readout over all N rows, `map_model` Linear.  Parameter names match the reference
(`conv_first.attention_0.w`, `conv_block.0.attention_1.a`, ...).

Quirk reproduced: the softmax runs over dim=1 of the broadcast [1,N,N] tensor, i.e.
over the ROW index i for every column j.  A column without edges (padded node, isolated real node)
is uniform 1/N over all N rows upstream and so adds (1/N) h_j to EVERY row of its graph; all padded
rows of a graph therefore share one value.  The kernels handle the edges; this module carries the
per-graph padded-row state and the correction vector with a few [G, F] torch ops.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import call, lib, ptr, stream_ptr, workspace
from .ops import CSR, READOUT_MAX, READOUT_SUM


class _GatAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, s1, s2, csr: CSR, heads: int, feat: int, slope: float):
        h = h.contiguous(); s1 = s1.contiguous(); s2 = s2.contiguous()
        n = h.size(0)
        mx = torch.empty(n, heads, dtype=torch.float32, device=h.device)
        zs = torch.empty(n, heads, dtype=torch.float32, device=h.device)
        hp = torch.empty_like(h)
        call("tsg_gat_fwd", ptr(csr.rowptr), ptr(csr.colidx), ptr(csr.t_rowptr), ptr(csr.t_colidx), ptr(h),
             ptr(s1), ptr(s2), n, heads, feat, float(slope), ptr(mx), ptr(zs), ptr(hp), stream_ptr())
        ctx.csr, ctx.heads, ctx.feat, ctx.slope = csr, heads, feat, float(slope)
        ctx.save_for_backward(h, s1, s2, mx, zs)
        return hp

    @staticmethod
    def backward(ctx, dhp):
        h, s1, s2, mx, zs = ctx.saved_tensors
        csr = ctx.csr
        dhp = dhp.contiguous()
        n, nnz = h.size(0), csr.colidx.numel()
        dh = torch.empty_like(h); ds1 = torch.empty_like(s1); ds2 = torch.empty_like(s2)
        wsb = lib.tsg_gat_bwd_workspace_bytes(nnz, ctx.heads)
        ws = workspace(wsb, h.device)
        call("tsg_gat_bwd", ptr(csr.rowptr), ptr(csr.eid), ptr(csr.t_rowptr), ptr(csr.t_colidx), ptr(csr.t_eid),
             ptr(h), ptr(s1), ptr(s2), ptr(mx), ptr(zs), ptr(dhp), n, nnz, ctx.heads, ctx.feat, ctx.slope,
             ptr(dh), ptr(ds1), ptr(ds2), ptr(ws), wsb, stream_ptr())
        return dh, ds1, ds2, None, None, None, None


class DGATHead(nn.Module):
    """Parameters of encoders_GAT.py:11-27: w [in,out], a [2*out,1], xavier-uniform gain 1.414."""

    def __init__(self, input_dim: int, output_dim: int, neg_input_slope: float = 0.2):
        super().__init__()
        self.input_dim, self.output_dim, self.slope = input_dim, output_dim, neg_input_slope
        self.w = nn.Parameter(torch.zeros(input_dim, output_dim))
        nn.init.xavier_uniform_(self.w.data, gain=1.414)
        self.a = nn.Parameter(torch.zeros(2 * output_dim, 1))
        nn.init.xavier_uniform_(self.a.data, gain=1.414)


class DGATLayer(nn.Module):
    def __init__(self, input_dim, output_dim, n_heads=4, concat=True, neg_input_slope=0.2):
        super().__init__()
        self.concat, self.n_heads, self.out, self.slope = concat, n_heads, output_dim, neg_input_slope
        for i in range(n_heads):
            self.add_module(f"attention_{i}", DGATHead(input_dim, output_dim, neg_input_slope))

    def heads(self) -> List[DGATHead]:
        return [getattr(self, f"attention_{i}") for i in range(self.n_heads)]

    def forward(self, x, x_pad, csr: CSR, graph_ptr, batch, iso, num_pad, max_nodes: int):
        """x [n, Fin] real rows, x_pad [G, Fin] the value shared by every padded row of each graph."""
        hs = self.heads()
        Hd, Fo = self.n_heads, self.out
        wcat = torch.cat([h.w for h in hs], dim=1)                              # [Fin, Hd*Fo]
        a1 = torch.stack([h.a[:Fo, 0] for h in hs]); a2 = torch.stack([h.a[Fo:, 0] for h in hs])
        h = ops.linear(x, wcat)                                                 # encoders_GAT.py:32
        # :35-36 in closed form: s1[i,hd] = h[i,hd,:] . a[:F], s2 = h . a[F:], as ONE tall-skinny product with the
        # block-diagonal [Hd*Fo, 2*Hd] matrix of the attention vectors (K3's GEMV-shaped kernels, fwd and bwd)
        eye = torch.eye(Hd, device=x.device, dtype=x.dtype)
        amat = torch.cat([(a1.unsqueeze(2) * eye.unsqueeze(1)).reshape(Hd * Fo, Hd),
                          (a2.unsqueeze(2) * eye.unsqueeze(1)).reshape(Hd * Fo, Hd)], dim=1)
        s12 = ops.linear(h, amat)
        s1, s2 = s12[:, :Hd].contiguous(), s12[:, Hd:].contiguous()
        raw = _GatAggregate.apply(h, s1, s2, csr, Hd, Fo, self.slope)           # :38-43 on the edges
        # columns without edges: uniform 1/N over all N rows (:39-41 with every entry masked)
        h_pad = x_pad @ wcat                                                    # [G, Hd*Fo]
        iso_sum = ops.readout(h * iso.view(-1, 1), graph_ptr, READOUT_SUM)      # sum of isolated real h_j
        corr = (iso_sum + num_pad.view(-1, 1) * h_pad) / float(max_nodes)
        hp, hp_pad = raw + ops.broadcast_rows(corr, graph_ptr, raw.size(0)), corr
        if self.concat:                                                         # :46-49, :75-76
            return F.elu(hp), F.elu(hp_pad)
        avg = hp.view(-1, Hd, Fo).mean(1); avg_pad = hp_pad.view(-1, Hd, Fo).mean(1)   # :78-83
        return F.elu(avg), F.elu(avg_pad)


class PackedGatEncoder(nn.Module):
    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers=2, num_heads=(2, 2),
                 final_dim="output_dim"):
        super().__init__()
        self.num_layers, self.final_dim = num_layers, final_dim
        nh = list(num_heads)
        self.conv_first = DGATLayer(input_dim, hidden_dim, nh[0], True)
        self.conv_block = nn.ModuleList([DGATLayer(hidden_dim * nh[i - 1], hidden_dim, nh[i], True)
                                         for i in range(1, num_layers - 1)]) if num_layers >= 3 else None
        self.conv_last = DGATLayer(hidden_dim * nh[-1], embedding_dim, nh[-1], False)
        self.pred_model = nn.Linear(embedding_dim, label_dim)
        self.map_model = nn.Linear(embedding_dim, embedding_dim)

    def layers(self):
        return [self.conv_first] + (list(self.conv_block) if self.conv_block is not None else []) + [self.conv_last]

    def readout(self, x, csr: CSR, graph_ptr, max_nodes: int):
        """encoders_GAT.py:175-189.  csr = RAW CSR (both orientations, with eids) of adj > 0."""
        G = graph_ptr.numel() - 1
        n_g = (graph_ptr[1:] - graph_ptr[:-1])
        batch = torch.repeat_interleave(torch.arange(G, device=x.device), n_g)
        num_pad = (max_nodes - n_g).to(torch.float32)
        iso = ((csr.t_rowptr[1:] - csr.t_rowptr[:-1]) == 0).to(torch.float32)
        x_pad = x.new_zeros(G, x.size(1))
        for layer in self.layers():
            x, x_pad = layer(x, x_pad, csr, graph_ptr, batch, iso, num_pad, max_nodes)
        out = ops.readout(x, graph_ptr, READOUT_MAX)
        return torch.where((num_pad > 0).view(-1, 1), torch.maximum(out, x_pad), out)

    def forward(self, x, csr: CSR, graph_ptr, max_nodes: int):
        r = self.readout(x, csr, graph_ptr, max_nodes)
        head = self.pred_model if self.final_dim != "output_dim" else self.map_model     # :191-198
        return r, head(r)
