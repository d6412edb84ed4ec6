"""Generates tests/golden/eigpool.npz by running the REAL reference preprocessing
(`/root/reference/Code/eigengcn/coarsen_pooling_with_last_eigen_padding.py::Graphs._coarserning_pooling_`)
in this container.  Stubs: `matplotlib`, `community` (unused by the function), and sklearn's
SpectralClustering is replaced by an object that returns GIVEN labels (the clustering's RNG is not part of the
path under test; cluster labels are an input of K11).  Run from the repo root:  python oracle/make_golden_eigpool.py"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "two-stage-gnn_b200"))
REF = "/root/reference/Code/eigengcn"


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "community"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import scipy.sparse
    if not hasattr(scipy.sparse, "csr"):                      # `scipy.sparse.csr.csr_matrix` (graph.py:136)
        scipy.sparse.csr = types.SimpleNamespace(csr_matrix=scipy.sparse.csr_matrix)
    elif not hasattr(scipy.sparse.csr, "csr_matrix"):
        scipy.sparse.csr.csr_matrix = scipy.sparse.csr_matrix
    import coarsen_pooling_with_last_eigen_padding as cp
    return cp


def main():
    cp = load_reference()
    from tsg import synth
    out = {}
    rng = np.random.default_rng(0)
    corpus = synth.make_corpus("PROTEINS", 6, seed=11)
    for g in range(6):
        n = corpus.num_nodes(g)
        e0, e1 = int(corpus.edge_ptr[g]), int(corpus.edge_ptr[g + 1])
        adj = np.zeros((n, n)); adj[corpus.row[e0:e1], corpus.col[e0:e1]] = 1.0
        k = max(1, n // 5)
        # labels: contiguous chunks of a random permutation -> every cluster has >= 2 nodes (the reference
        # returns -1 for singleton clusters, :175-176)
        perm = rng.permutation(n)
        labels = np.empty(n, np.int64)
        for c in range(k):
            labels[perm[c::k]] = c

        class FakeSC:
            def __init__(self, **kw): pass
            def fit(self, a): self.labels_ = labels
        cp.SpectralClustering = FakeSC
        gr = cp.Graphs(sp.csr_matrix(adj), [5])
        res, a_coarse, pms = gr._coarserning_pooling_(sp.csr_matrix(adj), 5, False)
        assert res == 1
        out[f"adj{g}"] = adj.astype(np.float32)
        out[f"labels{g}"] = labels
        out[f"coarse{g}"] = np.asarray(a_coarse.todense() if hasattr(a_coarse, "todense") else a_coarse, np.float64)
        out[f"pool{g}"] = np.stack([np.asarray(p.todense(), np.float64) for p in pms])      # [5, n, C]
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "eigpool.npz"), **out)
    print("wrote tests/golden/eigpool.npz", {k: v.shape for k, v in out.items() if k.endswith("0")})


if __name__ == "__main__":
    main()
