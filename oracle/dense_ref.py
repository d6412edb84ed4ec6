"""CPU ORACLE (test infrastructure, not product code) -- restatement of the reference's DENSE
modules: Code/sage+gat+diffpool/encoders.py, encoders_GAT.py and Code/eigengcn/encoders.py.

Every function cites the reference file:line it follows and keeps the reference's dense, zero-padded
[B, N, N] / [B, N, F] wire format and its quirks: raw 0/1 adjacency without self
loops, row L2-normalise after the bias, a FRESH BatchNorm1d(num_nodes) per call, max-readout over all
N rows including padding, GAT softmax over dim=1 of the broadcast [1,N,N] tensor.

PINNING: unlike the PyG half, these modules ARE importable from /root/reference in the build
container.  `oracle/make_golden.py` imports the real reference classes (with `.cuda()` neutralised),
runs them on seeded inputs and stores inputs / parameters / outputs / gradients under
`tests/golden/dense_*.npz`; `tests/test_oracle_dense.py` checks this restatement against those
fixtures.  The GPU box never sees /root/reference: only the fixtures travel.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------
# GraphConv + node-wise BN  (Code/sage+gat+diffpool/encoders.py:30-42,134-138;
#                            Code/eigengcn/encoders.py:28-41 is identical)
# ---------------------------------------------------------------------------------------------
def graph_conv(x: Tensor, adj: Tensor, weight: Tensor, bias: Optional[Tensor],
               add_self: bool = False, normalize_embedding: bool = True) -> Tensor:
    y = torch.matmul(adj, x)                         # encoders.py:33
    if add_self:
        y = y + x                                    # :34-35
    y = torch.matmul(y, weight)                      # :36
    if bias is not None:
        y = y + bias                                 # :37-38
    if normalize_embedding:
        y = F.normalize(y, p=2, dim=2)               # :39-40
    return y


def apply_bn(x: Tensor) -> Tensor:
    """encoders.py:134-138: `nn.BatchNorm1d(x.size(1))(x)` on [B, N, F] -- a brand-new module in
    training mode each call: channel = node index, statistics over (B, F), biased variance,
    eps 1e-5, gamma 1, beta 0."""
    mean = x.mean(dim=(0, 2), keepdim=True)
    var = x.var(dim=(0, 2), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + 1e-5)


def construct_mask(max_nodes: int, batch_num_nodes: Sequence[int]) -> Tensor:
    """encoders.py:121-132 -> [B, max_nodes, 1] of 0/1."""
    m = torch.zeros(len(batch_num_nodes), max_nodes)
    for i, n in enumerate(batch_num_nodes):
        m[i, :int(n)] = 1.0
    return m.unsqueeze(2)


def gcn_forward(x: Tensor, adj: Tensor, convs: List[dict], bn: bool = True,
                embedding_mask: Optional[Tensor] = None) -> Tensor:
    """encoders.py:140-167: conv_first -> ReLU -> BN, conv_block..., conv_last (no ReLU/BN),
    concat of every layer's output along features, times the mask."""
    outs = []
    for i, c in enumerate(convs):
        x = graph_conv(x, adj, c["weight"], c["bias"])
        if i < len(convs) - 1:
            x = F.relu(x)
            if bn:
                x = apply_bn(x)
        outs.append(x)
    xt = torch.cat(outs, dim=2)
    if embedding_mask is not None:
        xt = xt * embedding_mask
    return xt


def gcn_encoder_readout(x: Tensor, adj: Tensor, convs: List[dict], bn: bool = True) -> Tensor:
    """GcnEncoderGraph.forward up to `output` (encoders.py:169-203, concat=True, num_aggs=1):
    max over dim=1 (ALL N rows, unmasked) of every layer's output, concatenated."""
    outs = []
    for i, c in enumerate(convs):
        x = graph_conv(x, adj, c["weight"], c["bias"])
        if i < len(convs) - 1:
            x = F.relu(x)
            if bn:
                x = apply_bn(x)
        outs.append(torch.max(x, dim=1)[0])
    return torch.cat(outs, dim=1)


# ---------------------------------------------------------------------------------------------
# DGATHead / DGATLayer / DGATEncoderGraph  (Code/sage+gat+diffpool/encoders_GAT.py:29-49,70-84,175-198)
# ---------------------------------------------------------------------------------------------
def dgat_head(inp: Tensor, adj: Tensor, w: Tensor, a: Tensor, slope: float = 0.2,
              concat: bool = True) -> Tensor:
    """encoders_GAT.py:29-49 with the closed form e_ij = a[:F].h_i + a[F:].h_j (identical to the
    reference's materialised [N*N, 2F] product).  Only input[0] is used (:32); adj [1,N,N]
    broadcasts so softmax(dim=1) normalises over the ROW index i for every column j (:41-43)."""
    h = torch.mm(inp[0], w)
    Fo = w.size(1)
    s1 = h @ a[:Fo, 0]
    s2 = h @ a[Fo:, 0]
    e = F.leaky_relu(s1.view(-1, 1) + s2.view(1, -1), slope)
    att = torch.where(adj > 0, e, torch.full_like(e, -9e15))      # [1,N,N] after broadcast
    att = F.softmax(att, dim=1)
    hp = torch.matmul(att, h)
    return F.elu(hp) if concat else hp


def dgat_layer(x: Tensor, adj: Tensor, heads: List[dict], concat: bool) -> Tensor:
    """encoders_GAT.py:70-84 (dropout = 0)."""
    outs = [dgat_head(x, adj, hd["w"], hd["a"], concat=concat) for hd in heads]
    if concat:
        return torch.cat(outs, dim=2)
    s = outs[0]
    for o in outs[1:]:
        s = s + o
    return F.elu(s / len(outs))


def dgat_encoder_readout(x: Tensor, adj: Tensor, layers: List[List[dict]]) -> Tensor:
    """DGATEncoderGraph.forward up to the max-readout (encoders_GAT.py:175-189)."""
    for li, heads in enumerate(layers):
        x = dgat_layer(x, adj, heads, concat=li < len(layers) - 1)
    return torch.max(x, dim=1)[0]


# ---------------------------------------------------------------------------------------------
# SoftPoolingGcnEncoder (DiffPool)  (Code/sage+gat+diffpool/encoders.py:327-406)
# ---------------------------------------------------------------------------------------------
def soft_pool_readout(x: Tensor, adj: Tensor, batch_num_nodes, p: Dict, bn: bool = True) -> Tensor:
    """One pooling level (num_pooling=1): returns the concatenated readout `output` [B, 2*D]."""
    N = adj.size(1)
    mask = construct_mask(N, batch_num_nodes) if batch_num_nodes is not None else None
    z = gcn_forward(x, adj, p["conv"], bn, mask)                          # :350
    out0 = torch.max(z, dim=1)[0]                                         # :353
    za = gcn_forward(x, adj, p["assign_conv"], bn, mask)                  # :365 (x_a = x)
    s = F.softmax(F.linear(za, p["assign_pred.weight"], p["assign_pred.bias"]), dim=-1)   # :369
    if mask is not None:
        s = s * mask                                                      # :370-371
    xp = torch.matmul(s.transpose(1, 2), z)                               # :374
    ap = s.transpose(1, 2) @ adj @ s                                      # :375
    z2 = gcn_forward(xp, ap, p["conv_after"], bn, None)                   # :378
    out1 = torch.max(z2, dim=1)[0]                                        # :383
    return torch.cat([out0, out1], dim=1), dict(s=s, xp=xp, ap=ap, z=z)


def link_pred_loss(s: Tensor, adj: Tensor, batch_num_nodes, eps: float = 1e-7) -> Tensor:
    """The linkpred branch of SoftPoolingGcnEncoder.loss (encoders.py:416-440, adj_hop = 1) with the intended clamp at 1:
    s [B, N, K] masked assignment, adj [B, N, N]; entries outside the n x n block are dropped; normalised by sum n^2."""
    pred = torch.clamp(s @ s.transpose(1, 2), max=1.0)
    ll = -adj * torch.log(pred + eps) - (1 - adj) * torch.log(1 - pred + eps)
    mask = construct_mask(adj.size(1), batch_num_nodes).to(s.dtype)
    ll = ll * (mask @ mask.transpose(1, 2))
    return ll.sum() / float(sum(int(n) * int(n) for n in batch_num_nodes))


# ---------------------------------------------------------------------------------------------
# EigenPooling  (Code/eigengcn/encoders.py:396-417 Pool; :323-378 WavePoolingGcnEncoder.forward)
# ---------------------------------------------------------------------------------------------
def eigen_pool(x: Tensor, pool_matrices: List[Tensor]) -> Tensor:
    """Pool.forward: X_j = P_j^T X, concatenated over j on the feature axis."""
    res = [torch.matmul(pm.transpose(1, 2), x) for pm in pool_matrices]
    return torch.cat(res, dim=2) if len(res) > 1 else res[0]


def wave_readout(x: Tensor, adj: Tensor, adj_pooled: List[Tensor], batch_num_nodes,
                 batch_num_nodes_list, pool_matrices: List[List[Tensor]], p: Dict,
                 num_pool_matrix: int, num_pool_final_matrix: int, bn: bool = True) -> Tensor:
    """WavePoolingGcnEncoder.forward up to `output` (concat=True, mask=1, con_final=1)."""
    N = adj.size(1)
    mask = construct_mask(N, batch_num_nodes)
    outs = []
    z = gcn_forward(x, adj, p["conv"], bn, mask)                          # :333
    outs.append(torch.max(z, dim=1)[0])                                   # :337
    i = -1
    for i in range(len(adj_pooled)):
        z = eigen_pool(z, pool_matrices[i][:num_pool_matrix])             # :346-348
        m = construct_mask(N, batch_num_nodes_list[i])                    # :349-350
        z = gcn_forward(z, adj_pooled[i], p["conv_after"][i], bn, m)      # :353-356
        outs.append(torch.max(z, dim=1)[0])                               # :358-360
    if num_pool_final_matrix > 0:
        z = eigen_pool(z, pool_matrices[i + 1][:num_pool_final_matrix])   # :363-367
        outs.append(torch.max(z, dim=1)[0])
    return torch.cat(outs, dim=1)


def mlp(x: Tensor, layers: List[dict]) -> Tensor:
    """build_pred_layers: Linear (+ReLU between) (encoders.py:105-119)."""
    for i, l in enumerate(layers):
        x = F.linear(x, l["weight"], l["bias"])
        if i < len(layers) - 1:
            x = F.relu(x)
    return x
