"""CPU ORACLE (test infrastructure, not product code) -- PyG 1.6.3 operator restatement.

This file restates, in plain single-threaded-order torch CPU ops, the arithmetic of the
PyTorch-Geometric operators that `Code/sag` of the reference calls.  The arithmetic itself
lives in an UN-VENDORED third-party dependency: `torch-geometric`, pinned by the reference's
`README.md:20` as "1.16.3" (a typo; 1.6.3 is the torch-1.7-era release) which delegates to
`torch-scatter` 2.0.5.  Neither is installable here (no network), and the reference ships no
tests / golden vectors, so for this half of the oracle:

        *** PARITY UNPINNED against upstream PyG binaries ***

It is anchored instead on (i) the reference's own call sites (cited per function), (ii) the
published PyG 1.6.3 algorithm, (iii) a cross-check of GCN normalisation
against the reference's importable dense `GraphConv` on a pre-normalised adjacency
(`tests/test_oracle.py::test_gcn_cross_check_dense`), and (iv) brute-force pure-python loop
versions of every integer op (`*_loops` below) that the vectorised versions must equal.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  The product path (`two-stage-gnn_b200/tsg`) never does.

One deliberate deviation, documented: PyG writes `deg.pow(-0.5)`.  ATen's CPU kernels lower
pow(-0.5) / sqrt to ISA-dependent vector approximations (on this AVX512 host they differ from the
correctly rounded result by up to 2 ulp, e.g. deg=267) so they are not reproducible across hosts.
The oracle pins the IEEE form `1.0 / sqrt(deg)` computed with correctly rounded sqrt and divide
(`_ieee_rsqrt`); the difference is <= 2.4e-7 relative.  Passing float64 tensors runs the same
restatement in double precision (used by the tests to measure how well conditioned a quantity is).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# A.1.1  GCNConv   (reference call sites: Code/sag/network.py:19-23,34,38,42; layers.py:12,18)
# --------------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
                             fill_value: float = 1.0) -> Tuple[Tensor, Tensor]:
    """PyG 1.6.3 `utils.loop.add_remaining_self_loops`: drop existing self loops from the edge
    list, append one loop per node at the END (ids 0..N-1 in order); an existing loop keeps its
    weight (last one wins), every other node gets `fill_value`."""
    row, col = edge_index[0], edge_index[1]
    E = row.numel()
    w = torch.ones(E, dtype=torch.float32) if edge_weight is None else edge_weight
    mask = row != col
    loop_w = torch.full((num_nodes,), fill_value, dtype=w.dtype)
    inv = ~mask
    if bool(inv.any()):
        # sequential assignment => the last listed self-loop's weight wins
        for r, ww in zip(row[inv].tolist(), w[inv].tolist()):
            loop_w[r] = ww
    loop = torch.arange(num_nodes, dtype=row.dtype)
    row2 = torch.cat([row[mask], loop])
    col2 = torch.cat([col[mask], loop])
    w2 = torch.cat([w[mask], loop_w])
    return torch.stack([row2, col2]), w2


def gcn_norm(edge_index: Tensor, edge_weight: Optional[Tensor], num_nodes: int,
             dtype: torch.dtype = torch.float32) -> Tuple[Tensor, Tensor]:
    """PyG 1.6.3 `GCNConv.norm` / `gcn_norm` (improved=False, add_self_loops=True):
    deg = scatter_add(w, col); norm = deg^-1/2[row] * w * deg^-1/2[col]."""
    ei, w = add_remaining_self_loops(edge_index, edge_weight, num_nodes, 1.0)
    w = w.to(dtype)
    row, col = ei[0], ei[1]
    deg = torch.zeros(num_nodes, dtype=dtype).index_add_(0, col, w)
    dis = _ieee_rsqrt(deg)            # IEEE form of deg.pow(-0.5); see module docstring
    dis[dis == float("inf")] = 0.0
    norm = dis[row] * w * dis[col]    # evaluated left to right: (dis[row]*w)*dis[col]
    return ei, norm


def _ieee_rsqrt(deg: Tensor) -> Tensor:
    """1/sqrt(deg) with correctly rounded sqrt and divide (numpy uses the hardware IEEE
    instructions; ATen's vectorised CPU sqrt/rsqrt are ISA-dependent approximations that differ
    in the last bits between hosts, which would make the oracle irreproducible)."""
    import numpy as np
    d = deg.numpy()
    with np.errstate(divide="ignore"):
        out = (d.dtype.type(1.0) / np.sqrt(d)).astype(d.dtype)
    return torch.from_numpy(out)


def spmm_coo_edge_order(ei: Tensor, norm: Tensor, h: Tensor, num_nodes: int) -> Tensor:
    """out[col] += norm * h[row], accumulated sequentially in edge order (torch CPU index_add_
    on dim 0 is a sequential loop over the index vector), product rounded before the add."""
    msg = norm.view(-1, 1) * h[ei[0]]
    return torch.zeros(num_nodes, h.size(1), dtype=h.dtype).index_add_(0, ei[1], msg)


def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Optional[Tensor],
             edge_weight: Optional[Tensor] = None) -> Tensor:
    """GCNConv.forward with defaults (improved=False, cached=False, normalize=True)."""
    n = x.size(0)
    ei, norm = gcn_norm(edge_index, edge_weight, n, dtype=x.dtype)   # float64 = conditioning probe
    h = x @ weight
    out = spmm_coo_edge_order(ei, norm, h, n)
    if bias is not None:
        out = out + bias
    return out


def spmm_loops(ei: Tensor, norm: Tensor, h: Tensor, num_nodes: int) -> Tensor:
    """Brute-force fp32 loop version of `spmm_coo_edge_order` (small inputs only)."""
    import numpy as np
    out = np.zeros((num_nodes, h.size(1)), dtype=np.float32)
    hn, nn_ = h.numpy(), norm.numpy()
    for e in range(ei.size(1)):
        r, c = int(ei[0, e]), int(ei[1, e])
        out[c] = (out[c] + (nn_[e] * hn[r]).astype(np.float32)).astype(np.float32)
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------
# A.1.6  CSR construction contract (new; the reference stays COO)
# --------------------------------------------------------------------------------------
def csr_from_coo(ei: Tensor, val: Tensor, num_nodes: int, by: str = "dst"
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Stable counting sort of a COO list by destination (`by='dst'`, key=ei[1], neighbours =
    ei[0]) or by source (`by='src'`, the transposed operator).  Within a row the entries keep
    COO order, so a sequential per-row accumulation reproduces `index_add_` order bit for bit.
    Returns (rowptr[int32 N+1], colidx[int32 nnz], val[f32 nnz], eid[int32 nnz])."""
    key, other = (ei[1], ei[0]) if by == "dst" else (ei[0], ei[1])
    _, order = torch.sort(key, stable=True)
    counts = torch.bincount(key, minlength=num_nodes)
    rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(counts, 0)
    return (rowptr.to(torch.int32), other[order].to(torch.int32), val[order].clone(),
            order.to(torch.int32))


def gcn_csr(edge_index: Tensor, num_nodes: int, by: str = "dst"):
    """CSR of the self-loop-augmented, symmetrically normalised operator (K1's contract)."""
    ei, norm = gcn_norm(edge_index, None, num_nodes)
    return csr_from_coo(ei, norm, num_nodes, by)


def csr_from_coo_loops(ei: Tensor, val: Tensor, num_nodes: int, by: str = "dst"):
    key, other = (ei[1], ei[0]) if by == "dst" else (ei[0], ei[1])
    rows: List[List[int]] = [[] for _ in range(num_nodes)]
    for e in range(ei.size(1)):
        rows[int(key[e])].append(e)
    rowptr, colidx, v, eid = [0], [], [], []
    for r in range(num_nodes):
        for e in rows[r]:
            colidx.append(int(other[e])); v.append(float(val[e])); eid.append(e)
        rowptr.append(len(colidx))
    return (torch.tensor(rowptr, dtype=torch.int32), torch.tensor(colidx, dtype=torch.int32),
            torch.tensor(v, dtype=torch.float32), torch.tensor(eid, dtype=torch.int32))


def spmm_csr_sequential(rowptr: Tensor, colidx: Tensor, val: Tensor, h: Tensor) -> Tensor:
    """Row-sequential CSR SpMM in fp32 with rounded products (small inputs only)."""
    import numpy as np
    n = rowptr.numel() - 1
    out = np.zeros((n, h.size(1)), dtype=np.float32)
    hn, vn = h.numpy(), val.numpy()
    rp, ci = rowptr.tolist(), colidx.tolist()
    for r in range(n):
        acc = np.zeros(h.size(1), dtype=np.float32)
        for p in range(rp[r], rp[r + 1]):
            acc = (acc + (vn[p] * hn[ci[p]]).astype(np.float32)).astype(np.float32)
        out[r] = acc
    return torch.from_numpy(out)


# --------------------------------------------------------------------------------------
# A.1.2  topk   (reference call site: Code/sag/layers.py:20)
# --------------------------------------------------------------------------------------
def topk(x: Tensor, ratio: float, batch: Tensor) -> Tensor:
    """PyG 1.6.3 `topk_pool.topk` (min_score=None): per-graph descending sort of the score,
    keep the first k_g = ceil(ratio * n_g) (computed in fp32).  Tie rule frozen to STABLE
    (equal scores keep ascending node id), which is what torch's CPU sort does; NaN sorts
    first in descending order."""
    num_nodes = torch.bincount(batch) if batch.numel() else batch.new_zeros(0)
    batch_size = int(num_nodes.numel())
    if batch_size == 0:
        return batch.new_zeros(0)
    max_num_nodes = int(num_nodes.max())
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)[:-1]])
    index = torch.arange(batch.size(0), dtype=torch.long)
    index = (index - cum[batch]) + (batch * max_num_nodes)
    dense_x = x.new_full((batch_size * max_num_nodes,), torch.finfo(x.dtype).min)
    dense_x[index] = x
    dense_x = dense_x.view(batch_size, max_num_nodes)
    _, perm = dense_x.sort(dim=-1, descending=True, stable=True)
    perm = perm + cum.view(-1, 1)
    perm = perm.view(-1)
    k = (num_nodes.to(torch.float32) * ratio).ceil().to(torch.long)   # fp32 arithmetic
    mask = [torch.arange(int(k[i]), dtype=torch.long) + i * max_num_nodes
            for i in range(batch_size)]
    mask = torch.cat(mask, 0)
    return perm[mask]


def topk_loops(x: Tensor, ratio: float, batch: Tensor) -> Tensor:
    """Brute-force version: python sort with key (-score, index) per graph."""
    import numpy as np
    out: List[int] = []
    xs = x.tolist()
    b = batch.tolist()
    G = (max(b) + 1) if b else 0
    start = 0
    for g in range(G):
        n = b.count(g)
        ids = list(range(start, start + n))

        def key(i):
            v = xs[i]
            if math.isnan(v):
                return (0, 0.0, i)
            return (1, -v, i)
        ids.sort(key=key)
        k = int(np.ceil(np.float32(ratio) * np.float32(n)))
        out.extend(ids[:k])
        start += n
    return torch.tensor(out, dtype=torch.long)


# --------------------------------------------------------------------------------------
# A.1.3  filter_adj   (reference call site: Code/sag/layers.py:23)
# --------------------------------------------------------------------------------------
def filter_adj(edge_index: Tensor, edge_attr: Optional[Tensor], perm: Tensor,
               num_nodes: int) -> Tuple[Tensor, Optional[Tensor]]:
    mask = perm.new_full((num_nodes,), -1)
    i = torch.arange(perm.size(0), dtype=torch.long)
    mask[perm] = i
    row, col = edge_index[0], edge_index[1]
    row, col = mask[row], mask[col]
    keep = (row >= 0) & (col >= 0)
    row, col = row[keep], col[keep]
    if edge_attr is not None:
        edge_attr = edge_attr[keep]
    return torch.stack([row, col], dim=0), edge_attr


def filter_adj_loops(edge_index: Tensor, perm: Tensor, num_nodes: int) -> Tensor:
    m = [-1] * num_nodes
    for i, p in enumerate(perm.tolist()):
        m[p] = i
    r_out, c_out = [], []
    for r, c in zip(edge_index[0].tolist(), edge_index[1].tolist()):
        if m[r] >= 0 and m[c] >= 0:
            r_out.append(m[r]); c_out.append(m[c])
    return torch.tensor([r_out, c_out], dtype=torch.long).view(2, -1)


# --------------------------------------------------------------------------------------
# A.1.4  readouts   (reference call sites: Code/sag/network.py:36,40,44)
# --------------------------------------------------------------------------------------
def global_mean_pool(x: Tensor, batch: Tensor, size: Optional[int] = None) -> Tensor:
    size = int(batch.max()) + 1 if size is None else size
    s = torch.zeros(size, x.size(1), dtype=x.dtype).index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=size).clamp(min=1).to(x.dtype).view(-1, 1)
    return s / cnt


def global_max_pool(x: Tensor, batch: Tensor, size: Optional[int] = None) -> Tensor:
    """segment max; empty segment -> 0; gradient goes to the FIRST arg-max row (torch-scatter
    CPU semantics)."""
    size = int(batch.max()) + 1 if size is None else size
    cnt = torch.bincount(batch, minlength=size)
    ptr = torch.cat([cnt.new_zeros(1), cnt.cumsum(0)])
    outs = []
    for g in range(size):
        lo, hi = int(ptr[g]), int(ptr[g + 1])
        if hi == lo:
            outs.append(x.new_zeros(x.size(1)))
        else:
            seg = x[lo:hi]
            # first argmax per column, routed through gather so autograd hits one row
            am = _first_argmax(seg)
            outs.append(seg.gather(0, am.view(1, -1)).view(-1))
    return torch.stack(outs, 0)


def _first_argmax(seg: Tensor) -> Tensor:
    mx = seg.max(dim=0).values
    is_max = seg == mx.view(1, -1)
    idx = torch.arange(seg.size(0)).view(-1, 1).expand_as(seg)
    big = seg.size(0)
    return torch.where(is_max, idx, torch.full_like(idx, big)).min(dim=0).values


# --------------------------------------------------------------------------------------
# A.1.5  Batch.from_data_list   (reference: via DataLoader, Code/sag/train.py:181)
# --------------------------------------------------------------------------------------
def batch_from_data_list(xs: Sequence[Tensor], eis: Sequence[Tensor], ys: Sequence[Tensor]):
    off, ei_out, batch = 0, [], []
    for g, (x, ei) in enumerate(zip(xs, eis)):
        ei_out.append(ei + off)
        batch.append(torch.full((x.size(0),), g, dtype=torch.long))
        off += x.size(0)
    return (torch.cat(list(xs), 0), torch.cat(ei_out, 1), torch.cat(batch, 0),
            torch.cat([y.view(-1) for y in ys], 0))


# --------------------------------------------------------------------------------------
# SAGPool layer and Net (restating Code/sag/layers.py:14-26 and network.py:30-53 with the
# `batch` vector the reference's own commented-out line network.py:31 would have passed).
# --------------------------------------------------------------------------------------
def sag_pool(x: Tensor, edge_index: Tensor, batch: Tensor, w_score: Tensor, b_score: Tensor,
             ratio: float):
    score = gcn_conv(x, edge_index, w_score, b_score).squeeze(-1)          # layers.py:18
    perm = topk(score, ratio, batch)                                        # layers.py:20
    x = x[perm] * torch.tanh(score[perm]).view(-1, 1)                       # layers.py:21
    batch = batch[perm]                                                     # layers.py:22
    edge_index, _ = filter_adj(edge_index, None, perm, num_nodes=score.size(0))   # layers.py:23
    return x, edge_index, batch, perm, score


def sag_net_forward(params: dict, x: Tensor, edge_index: Tensor, batch: Tensor, ratio: float,
                    dropout_mask: Optional[Tensor] = None, return_aux: bool = False):
    """Net.forward (network.py:30-53).  `params` uses the reference's state_dict keys
    (`conv1.weight` [in,out], `pool1.score_layer.weight`, `lin1.weight` [out,in], ...).
    Dropout (network.py:49) is replaced by an injected keep-mask already scaled by 1/(1-p)
    (or None = eval mode) so runs are reproducible."""
    G = int(batch.max()) + 1
    aux = {"perm": [], "edge_index": [], "score": []}
    xs = []
    for lvl in (1, 2, 3):
        x = F.relu(gcn_conv(x, edge_index, params[f"conv{lvl}.weight"], params[f"conv{lvl}.bias"]))
        x, edge_index, batch, perm, score = sag_pool(
            x, edge_index, batch, params[f"pool{lvl}.score_layer.weight"],
            params[f"pool{lvl}.score_layer.bias"], ratio)
        aux["perm"].append(perm); aux["edge_index"].append(edge_index); aux["score"].append(score)
        xs.append(torch.cat([global_max_pool(x, batch, G), global_mean_pool(x, batch, G)], dim=1))
    x = xs[0] + xs[1] + xs[2]                                               # network.py:46
    x = F.relu(F.linear(x, params["lin1.weight"], params["lin1.bias"]))
    if dropout_mask is not None:
        x = x * dropout_mask
    x = F.relu(F.linear(x, params["lin2.weight"], params["lin2.bias"]))
    x = F.log_softmax(F.linear(x, params["lin3.weight"], params["lin3.bias"]), dim=-1)
    return (x, aux) if return_aux else x


# --------------------------------------------------------------------------------------
# Triplet distance + margin loss (Code/sag/tripletnet.py:21-22; train_triplet.py:196,208-211)
# --------------------------------------------------------------------------------------
def pairwise_distance(a: Tensor, b: Tensor, eps: float = 1e-6) -> Tensor:
    """F.pairwise_distance(a, b, 2): || a - b + eps ||_2 (eps added to the difference)."""
    return torch.sqrt(((a - b + eps) ** 2).sum(dim=-1))


def triplet_margin_loss(ea: Tensor, ep: Tensor, en: Tensor, alpha: float):
    """MarginRankingLoss(margin=alpha)(d_p, d_n, target=-1) = mean(max(0, (d_p - d_n) + alpha))."""
    dp, dn = pairwise_distance(ea, ep), pairwise_distance(ea, en)
    return torch.clamp((dp - dn) + alpha, min=0).mean(), dp, dn


def init_sag_params(num_features: int, nhid: int, num_classes: int, seed: int = 777) -> dict:
    """Parameters with the reference's initialisers (PyG glorot-uniform for GCNConv weight
    [in,out], zero bias; torch.nn.Linear default init), drawn under torch.manual_seed(seed)."""
    g = torch.Generator().manual_seed(seed)

    def glorot(i, o):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(i, o, generator=g) * 2 - 1) * a

    def linear(i, o):
        bound = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound     # kaiming_uniform(a=sqrt(5))
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        return w, b

    p = {}
    dims = [(num_features, nhid), (nhid, nhid), (nhid, nhid)]
    for lvl, (i, o) in enumerate(dims, start=1):
        p[f"conv{lvl}.weight"] = glorot(i, o); p[f"conv{lvl}.bias"] = torch.zeros(o)
        p[f"pool{lvl}.score_layer.weight"] = glorot(nhid, 1)
        p[f"pool{lvl}.score_layer.bias"] = torch.zeros(1)
    p["lin1.weight"], p["lin1.bias"] = linear(nhid * 2, nhid)
    p["lin2.weight"], p["lin2.bias"] = linear(nhid, nhid // 2)
    p["lin3.weight"], p["lin3.bias"] = linear(nhid // 2, num_classes)
    return p
