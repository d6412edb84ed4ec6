#!/usr/bin/env python
"""bench_configs.py -- the OTHER BASELINE.json configs end to end on one B200 (bench.py is config 2, the
headline).  One JSON line per config: a 2stg triplet training step (forward of every graph of T triplets
in ONE packed batch, MarginRankingLoss, backward, Adam) or, for config 1, the original cross-entropy step.

  1  GraphSAGE/base  GcnEncoderGraph(32,32,32,2,L=2,bn)            PROTEINS-shape, cross-entropy
  3  GAT 2stg+       DGATEncoderGraph(32,32,32,2,L=3,heads [2,2])  JAN.Y-shape, edge-softmax path (K4)
  4  DiffPool 2stg   SoftPoolingGcnEncoder(N=1000,...,K=100)       DD-shape, tcgen05 contraction (K7)
  5  EigenGCN 2stg+  WavePoolingGcnEncoder(89,32,32,2,L=2,pool [10]) DD-shape, eigen pooling (K8)

Synthetic graphs of the named shapes (tsg.synth, seed 777), features per SURVEY 8d (shared N(0,4) table
for the dense directories, one-hot node labels for eigengcn), random-init weights.  Not a driver
contract: evidence that every config trains on the GPU path, with its throughput.

  python bench_configs.py [--configs 1,3,4,5] [--steps 10] [--warmup 3] [--graphs N]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "two-stage-gnn_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graphs", type=int, default=0, help="corpus size override (0 = the README row's count)")
    return ap.parse_args()


def timed(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(steps):
        last = fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, last


def triplet_batch(corpus, seed):
    from tsg import synth
    T = corpus.num_graphs
    trip = synth.sample_triplets(corpus.y, T, seed=seed)
    ids = np.concatenate([trip[:, 0], trip[:, 1], trip[:, 2]])
    tidx = np.stack([np.arange(T), T + np.arange(T), 2 * T + np.arange(T)], 1).astype(np.int64)
    return ids, tidx


def dense_inputs(corpus, ids, dev, feat_dim=32, max_nodes=1000, want_eid=False):
    """packed x (rows of the shared feature table, Code/sage+gat+diffpool/train_triplet.py:379-385), the
    RAW 0/1 CSR of the packed adjacency, graph offsets and the has-padded-rows flags."""
    from tsg import ops, synth
    sel = synth.select(corpus, ids)
    n = np.diff(sel.node_ptr)
    table = torch.from_numpy(np.random.default_rng(777).normal(0.0, 2.0, (max_nodes, feat_dim)).astype(np.float32))
    local = np.concatenate([np.arange(k) for k in n])
    x = table[torch.from_numpy(local)].to(dev)
    pk = synth.pack(sel, one_hot=False)
    ei = torch.from_numpy(pk["edge_index"]).to(dev)
    N = int(sel.node_ptr[-1])
    csr = ops.build_csr(ops.EdgeList.from_edge_index(ei), N, mode=ops.CSR_RAW, want_eid=want_eid)
    gptr = torch.from_numpy(sel.node_ptr).to(dev)
    has_pad = torch.from_numpy(n < max_nodes).to(dev)
    return x, csr, gptr, has_pad, N, int(ei.size(1))


def line(cfg, name, model_desc, graphs, ms, N, E, loss, extra=None):
    d = {"config": cfg, "metric": "train graphs/sec (fwd+bwd)", "value": graphs / (ms / 1e3), "unit": "graphs/s",
         "ms_per_step": ms, "graphs_per_step": graphs, "nodes_per_step": N, "directed_edges_per_step": E,
         "workload": name, "model": model_desc, "loss": float(loss), "n_gpus": 1, "dtype": "f32", "data": "synthetic"}
    if extra:
        d.update(extra)
    assert np.isfinite(d["loss"]), "non-finite loss"
    print(json.dumps(d), flush=True)


def config1(a, dev):
    from tsg import dense, synth
    corpus = synth.make_corpus("PROTEINS", a.graphs or 1113, seed=777)
    ids = np.arange(corpus.num_graphs)
    x, csr, gptr, has_pad, N, E = dense_inputs(corpus, ids, dev, max_nodes=1000)
    y = torch.from_numpy(corpus.y).to(dev)
    torch.manual_seed(777)
    model = dense.PackedGcnEncoder(32, 32, 32, 2, 2, bn=True, final_dim="number_classes").to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step(i):
        _, logits = model(x, csr, gptr, has_pad)
        loss = torch.nn.functional.cross_entropy(logits, y)
        opt.zero_grad(set_to_none=True); loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 2.0)          # train.py clip 2.0
        opt.step()
        return loss.detach()
    ms, loss = timed(step, a.steps, a.warmup)
    line(1, "GraphSAGE original setting, PROTEINS-shape, one packed batch of the whole corpus, cross-entropy",
         "GcnEncoderGraph(32,32,32,2,L=2,bn,final_dim=number_classes)", corpus.num_graphs, ms, N, E, loss)


def config3(a, dev):
    from tsg import gat, ops, synth
    corpus = synth.make_corpus("JANY", a.graphs or 744, seed=777)
    ids, tidx = triplet_batch(corpus, 0)
    x, csr, gptr, has_pad, N, E = dense_inputs(corpus, ids, dev, want_eid=True)
    tr = torch.from_numpy(tidx).to(dev)
    torch.manual_seed(777)
    model = gat.PackedGatEncoder(32, 32, 32, 2, num_layers=3, num_heads=[2, 2, 2]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step(i):
        _, emb = model(x, csr, gptr, 1000)
        loss, _, _ = ops.triplet_loss(emb, tr, 1.5)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        return loss.detach()
    ms, loss = timed(step, a.steps, a.warmup)
    line(3, f"GAT 2stg+ stage-1 triplet step, JAN.Y-shape, 3x{corpus.num_graphs} graphs packed", "DGATEncoderGraph(32,32,32,2,L=3,heads 2)",
         ids.shape[0], ms, N, E, loss)


def config4(a, dev):
    from tsg import diffpool, ops, synth
    corpus = synth.make_corpus("DD", a.graphs or 1168, seed=777)
    ids, tidx = triplet_batch(corpus, 0)
    x, csr, gptr, has_pad, N, E = dense_inputs(corpus, ids, dev)
    tr = torch.from_numpy(tidx).to(dev)
    torch.manual_seed(777)
    model = diffpool.PackedSoftPoolEncoder(1000, 32, 32, 32, 2, 3, 32, assign_ratio=0.1).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step(i):
        _, emb = model(x, csr, gptr, has_pad)
        loss, _, _ = ops.triplet_loss(emb, tr, 1.5)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        return loss.detach()
    ms, loss = timed(step, a.steps, a.warmup)
    line(4, f"DiffPool 2stg triplet step, DD-shape, 3x{corpus.num_graphs} graphs packed, K=100 clusters, tcgen05 S^T[Z|AS]",
         "SoftPoolingGcnEncoder(N=1000,32,32,32,2,L=3,assign_ratio=0.1)", ids.shape[0], ms, N, E, loss,
         {"tensor_cores": bool(ops.USE_TCGEN05)})


def config5(a, dev):
    from tsg import dense, eigen_synth, ops, synth
    corpus = synth.make_corpus("DD", a.graphs or 1168, seed=777)
    opnd = eigen_synth.make_operands(corpus, pool_size=10, num_pool_matrix=1, num_pool_final_matrix=1)
    ids, tidx = triplet_batch(corpus, 0)
    sel = synth.select(corpus, ids)
    pk = synth.pack(sel)
    x = torch.from_numpy(pk["x"]).to(dev)
    ei = torch.from_numpy(pk["edge_index"]).to(dev)
    N, E = int(sel.node_ptr[-1]), int(ei.size(1))
    csr_adj = ops.build_csr(ops.EdgeList.from_edge_index(ei), N, mode=ops.CSR_RAW)
    po = eigen_synth.pack_operands(corpus, opnd, ids)
    NC, G = int(po["cluster_ptr"][-1]), ids.shape[0]
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)

    def rect(src, dst, w, n_out, n_in):
        return dense.build_rect_csr(ops.EdgeList(t(src), t(dst), int(src.shape[0])), t(w), n_out, n_in)
    # level-1 operands on the GPU (K11): cluster labels in, P_0^T and the coarsened adjacency out
    from tsg import eigenpool
    cl = t(po["pool"][0][1].astype(np.int32))                      # global cluster id of every packed node
    el_adj = ops.EdgeList.from_edge_index(ei)
    eigenpool.build(csr_adj, el_adj, cl, NC, 1)                     # warm-up
    torch.cuda.synchronize()
    s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_.record(); built = eigenpool.build(csr_adj, el_adj, cl, NC, 1); e_.record()
    torch.cuda.synchronize()
    build_ms = s_.elapsed_time(e_)
    assert int(built["status"].item()) == 0
    pool0 = [built["pool"][0]]
    coarse = built["coarse"]
    final = [rect(*po["final"][0], G, NC)]
    gptr, cptr = t(po["node_ptr"]), t(po["cluster_ptr"])
    fptr = torch.arange(G + 1, device=dev, dtype=torch.int64)
    tr = torch.from_numpy(tidx).to(dev)
    torch.manual_seed(777)
    model = dense.PackedWaveEncoder(corpus.num_node_labels, 32, 32, 2, 2, num_pool_matrix=1, num_pool_final_matrix=1,
                                    pool_sizes=[10], pred_hidden_dims=[50]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step(i):
        emb = model(x, csr_adj, gptr, [pool0, final], [coarse], [cptr], fptr)
        loss, _, _ = ops.triplet_loss(emb, tr, 1.5)
        opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
        return loss.detach()
    ms, loss = timed(step, a.steps, a.warmup)
    line(5, f"EigenGCN 2stg+ stage-1 triplet step, DD-shape, 3x{corpus.num_graphs} graphs packed, pool_sizes [10] (BFS-chunk clusters)",
         "WavePoolingGcnEncoder(89,32,32,2,L=2,num_pool_matrix=1,num_pool_final_matrix=1,pred_hidden [50])",
         G, ms, N, E, loss, {"clusters_per_step": NC, "eigpool_build_ms": build_ms,
                             "eigpool_clusters_per_s": NC / (build_ms / 1e3),
                             "eigpool_note": "K11: per-cluster Laplacian eigendecomposition + coarsened adjacency of the "
                                             "whole packed batch on the GPU (cluster labels from the host BFS-chunk stand-in)"})


def main():
    a = parse()
    if not torch.cuda.is_available():
        raise SystemExit("bench_configs.py needs a CUDA device: tsg has no CPU path")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    fns = {"1": config1, "3": config3, "4": config4, "5": config5}
    for c in a.configs.split(","):
        fns[c.strip()](a, dev)


if __name__ == "__main__":
    main()
