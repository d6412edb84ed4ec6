"""EigenPooling operands on the GPU (K11).

`build(...)` turns (adjacency, cluster labels) of a packed batch into what `tsg.dense.PackedWaveEncoder`
consumes: the pooling operators P_j^T as rectangular CSRs (forward = segment-weighted sum, backward = gather)
and the coarsened adjacency as a RAW CSR -- the device-side replacement of the per-graph numpy/scipy work in
Code/eigengcn/coarsen_pooling_with_last_eigen_padding.py:121-182.  Cluster labels are an input (the reference
uses sklearn SpectralClustering; `tsg.eigen_synth` uses BFS chunks for synthetic corpora)."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops
from ._lib import call, lib, ptr, stream_ptr, workspace
from .ops import CSR, CSR_RAW, EdgeList

EIG_MAX = 32


def cluster_members(cluster_of: torch.Tensor, num_clusters: int):
    """(member_ptr int32 [C+1], member int32 [N]): nodes of every cluster in ascending node order
    (K1 RAW counting sort of the (node -> cluster) list; stable, so node order is kept)."""
    n = cluster_of.numel()
    nodes = torch.arange(n, device=cluster_of.device, dtype=torch.int64)
    el = EdgeList(nodes, cluster_of.to(torch.int64), n)
    csr = ops.build_csr(el, max(n, num_clusters), mode=CSR_RAW, transposed=False)
    return csr.rowptr[:num_clusters + 1], csr.colidx


def build(csr_adj: CSR, edges: EdgeList, cluster_of: torch.Tensor, num_clusters: int, num_vectors: int,
          edge_weight: Optional[torch.Tensor] = None, want_eigvals: bool = False, check: bool = False):
    """csr_adj / edges: the packed (symmetric) adjacency as RAW CSR and as the COO list it was built from;
    cluster_of int32 [N] global cluster ids.  Returns dict(pool=[CSR]*num_vectors, coarse=CSR, eigvals, status)."""
    dev = cluster_of.device
    n, C = cluster_of.numel(), int(num_clusters)
    cluster_of = cluster_of.to(torch.int32).contiguous()
    member_ptr, member = cluster_members(cluster_of, C)
    pool_val = torch.zeros(num_vectors, n, dtype=torch.float32, device=dev)
    pool_val_nodes = torch.zeros(num_vectors, n, dtype=torch.float32, device=dev)
    eigvals = torch.empty(C, EIG_MAX, dtype=torch.float32, device=dev) if want_eigvals else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    call("tsg_eigpool_build", ptr(csr_adj.rowptr), ptr(csr_adj.colidx), ptr(csr_adj.val), ptr(cluster_of),
         ptr(member_ptr), ptr(member), n, C, num_vectors, ptr(pool_val), ptr(pool_val_nodes), ptr(eigvals), ptr(status),
         stream_ptr())
    if check and int(status.item()) != 0:
        raise RuntimeError(f"tsg.eigenpool: a cluster has more than {EIG_MAX} nodes (host fallback needed)")
    node_rowptr = torch.arange(n + 1, device=dev, dtype=torch.int32)
    pool: List[CSR] = []
    for j in range(num_vectors):
        pool.append(CSR(member_ptr, member, pool_val[j], None, node_rowptr, cluster_of, pool_val_nodes[j], None, C))
    # coarsened adjacency: inter-cluster edges relabelled, duplicates summed by the SpMM
    E = edges.cap
    orow = torch.empty(max(E, 1), dtype=torch.int64, device=dev)
    ocol = torch.empty(max(E, 1), dtype=torch.int64, device=dev)
    ow = torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    cnt = torch.empty(1, dtype=torch.int64, device=dev)
    wsb = lib.tsg_coarsen_edges_workspace_bytes(E)
    ws = workspace(wsb, dev)
    if edges.count is not None:
        raise RuntimeError("tsg.eigenpool: the adjacency edge list must have a host-known length")
    call("tsg_coarsen_edges", ptr(edges.row), ptr(edges.col), ptr(edge_weight), E, ptr(cluster_of), ptr(orow), ptr(ocol),
         ptr(ow), ptr(cnt), ptr(ws), wsb, stream_ptr())
    coarse = ops.build_csr(EdgeList(orow, ocol, E, cnt), C, mode=CSR_RAW, edge_weight=ow)
    return dict(pool=pool, coarse=coarse, coarse_coo=(orow, ocol, ow, cnt), eigvals=eigvals, status=status,
                member_ptr=member_ptr, member=member)
