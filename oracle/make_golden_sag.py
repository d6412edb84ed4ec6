"""Generate tests/golden/sag_glue.npz by running the UNMODIFIED reference glue -- /root/reference/Code/sag/network.py
`Net` and layers.py `SAGPool` -- on CPU, in this container, on top of a `torch_geometric` package whose operators are
the oracle (oracle/pyg_oracle_shim).  The fixture pins everything ABOVE the operator boundary to the reference's own
code: which conv feeds which pool, `.squeeze()` / `.view(-1, 1)`, ranking on the raw (un-tanh'd) score, the
[gmp || gap] order, x1 + x2 + x3, the lin1/lin2/lin3/log_softmax head.  The operators underneath stay the restatement
of PyG 1.6.3 ("glue pinned, ops unpinned").

    python oracle/make_golden_sag.py          # rewrites tests/golden/sag_glue.npz

`Net.forward` hard-codes batch=None (network.py:32): every forward is ONE graph, exactly what the scripts run.  The
packed path must reproduce each graph's row.  Stored per graph g: node labels (x = onehot), edge_index, the perm and
the filtered edge_index of each pooling level (forward hooks on pool1..3), the embedding (log-softmax vector), and the
gradients of sum_g <emb_g, cot_g> accumulated over all graphs.  Graphs: DD-shape and PROTEINS-shape synthetic graphs
plus hand-made edge cases (the smallest graph the glue survives, a path, a star).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/sag"
OUT = os.path.join(ROOT, "tests", "golden", "sag_glue.npz")
for p in (os.path.join(ROOT, "oracle", "pyg_oracle_shim"), ROOT, os.path.join(ROOT, "two-stage-gnn_b200")):
    sys.path.insert(0, p)


def graphs():
    from tsg import synth
    out = []
    dd = synth.make_corpus("DD", 6, seed=4242)
    for g in range(6):
        pk = synth.pack(dd, [g])
        out.append((dd.node_label[dd.node_ptr[g]:dd.node_ptr[g + 1]].astype(np.int64), pk["edge_index"]))
    pr = synth.make_corpus("PROTEINS", 5, seed=99)
    rng = np.random.default_rng(5)
    for g in range(5):
        pk = synth.pack(pr, [g])
        n = pr.num_nodes(g)
        out.append((rng.integers(0, 89, n).astype(np.int64), pk["edge_index"]))       # 89 labels: same Net
    # edge cases: the smallest graph the reference glue survives (5 nodes: 5 -> 3 -> 2 -> 1; with fewer, a pooled
    # level has ONE node, `score.squeeze()` is 0-d and layers.py:21 raises IndexError upstream); path of 7; star of 9
    # (distinct leaf labels: leaves with EQUAL labels have mathematically equal scores and the reference's own result
    # then depends on the GEMM blocking of the batch it happens to be in -- exact ties are an operator-level test)
    c5 = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 0]], np.int64)
    c5 = np.concatenate([c5, c5[:, ::-1]]); c5 = c5[np.lexsort((c5[:, 1], c5[:, 0]))].T
    out.append((np.array([3, 1, 4, 1, 5], np.int64), c5))
    path = np.array([[i, i + 1] for i in range(6)] + [[i + 1, i] for i in range(6)], np.int64)
    path = path[np.lexsort((path[:, 1], path[:, 0]))].T
    out.append((np.arange(7, dtype=np.int64), path))
    star = np.array([[0, i] for i in range(1, 9)] + [[i, 0] for i in range(1, 9)], np.int64)
    star = star[np.lexsort((star[:, 1], star[:, 0]))].T
    out.append((np.array([1, 2, 3, 5, 8, 13, 21, 34, 55], np.int64), star))
    return out


def main():
    sys.path.insert(0, REF)
    for m in ("layers", "network", "torch_geometric"):
        sys.modules.pop(m, None)
    import network                              # the reference's file, unmodified
    from oracle import pyg_ref as R
    from torch_geometric.data import Data
    assert network.__file__.startswith(REF), network.__file__
    F, nhid, C = 89, 32, 32
    net = network.Net(F, nhid, C, 0.5, 0.5)
    params = R.init_sag_params(F, nhid, C, seed=777)
    # non-zero conv / score biases so the bias paths are exercised (the initialiser leaves them at zero)
    g = torch.Generator().manual_seed(3)
    for k in params:
        if k.endswith(".bias") and (k.startswith("conv") or k.startswith("pool")):
            params[k] = torch.randn(params[k].shape, generator=g) * 0.1
    missing = net.load_state_dict(params, strict=True)
    net.eval()                                  # dropout off (network.py:49)
    cap = {}
    for lvl, pool in enumerate((net.pool1, net.pool2, net.pool3)):
        pool.register_forward_hook(lambda m, i, o, lvl=lvl: cap.__setitem__(lvl, (o[4].clone(), o[1].clone())))
    out = {"dims": np.array([F, nhid, C])}
    gs = graphs()
    cot = torch.randn(len(gs), C, generator=g)
    total = 0
    for gi, (lab, ei) in enumerate(gs):
        x = torch.zeros(lab.shape[0], F); x[torch.arange(lab.shape[0]), torch.from_numpy(lab)] = 1.0
        emb = net(Data(x, torch.from_numpy(ei), None))
        assert emb.shape == (1, C)
        total = total + (emb[0] * cot[gi]).sum()
        out[f"g{gi}/label"] = lab.astype(np.int32); out[f"g{gi}/edge_index"] = ei
        out[f"g{gi}/emb"] = emb[0].detach().numpy()
        for lvl in range(3):
            out[f"g{gi}/perm{lvl}"] = cap[lvl][0].numpy(); out[f"g{gi}/ei{lvl}"] = cap[lvl][1].numpy()
    total.backward()
    for k, p in net.named_parameters():
        out["param/" + k] = p.detach().numpy(); out["grad/" + k] = p.grad.numpy()
    out["cot"] = cot.numpy(); out["num_graphs"] = np.array(len(gs))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(gs), "graphs")


if __name__ == "__main__":
    main()
