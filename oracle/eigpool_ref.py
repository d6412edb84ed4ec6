"""ORACLE (test infrastructure only) -- numpy restatement of the post-clustering half of the reference's
EigenPooling preprocessing, `/root/reference/Code/eigengcn/coarsen_pooling_with_last_eigen_padding.py`:

  * per cluster: L = D - W of the induced subgraph (graph.laplacian(adj, normalized=False), graph.py:116-126,
    d = W.sum(axis=0)), `lamb, U = np.linalg.eigh(L)` (graph.fourier, graph.py:147-163), column j of the pooling
    matrix P_j gets U[:, j] on the cluster's rows, negated when U[0, j] < 0 (:163-168), and U[:, size-1]
    (same sign rule) when j >= size (:169-173);
  * A_int keeps intra-cluster entries, A_ext = A - A_int, A_coarsened = Omega^T A_ext Omega (:135-149).

Cluster labels are an input (upstream: sklearn SpectralClustering, :125-126, not reproducible without its RNG).
PARITY: pinned only against numpy's LAPACK `eigh` as installed here (the reference ships no fixtures for this
step); eigenvectors inside a degenerate eigenspace are compared through the eigenspace projector."""
from __future__ import annotations

from typing import List

import numpy as np


def cluster_eigvecs(adj: np.ndarray):
    """(eigenvalues ascending, eigenvectors as columns) of the unnormalised Laplacian of a dense symmetric W."""
    d = adj.sum(axis=0)
    lap = np.diag(d) - adj
    lamb, u = np.linalg.eigh(lap.astype(np.float64))
    return lamb, u


def pooling_matrices(adj: np.ndarray, clusters: List[List[int]], num_vectors: int):
    """dense P_j [n, C] for j < num_vectors, and A_coarsened [C, C]."""
    n, C = adj.shape[0], len(clusters)
    P = [np.zeros((n, C), np.float64) for _ in range(num_vectors)]
    label = np.full(n, -1, np.int64)
    for c, nodes in enumerate(clusters):
        label[np.asarray(nodes)] = c
        sub = adj[np.ix_(nodes, nodes)]
        _, u = cluster_eigvecs(sub)
        size = len(nodes)
        for j in range(num_vectors):
            col = u[:, min(j, size - 1)]
            P[j][np.asarray(nodes), c] = -col if col[0] < 0 else col
    a_ext = adj * (label[:, None] != label[None, :])
    omega = np.zeros((n, C), np.float64); omega[np.arange(n), label] = 1.0
    return P, omega.T @ a_ext @ omega
