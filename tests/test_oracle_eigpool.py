"""oracle/eigpool_ref.py pinned against tests/golden/eigpool.npz, which oracle/make_golden_eigpool.py produced by
running the REAL reference `_coarserning_pooling_` (SpectralClustering replaced by given labels)."""
import os

import numpy as np

from oracle import eigpool_ref as E

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eigpool.npz")


def test_restatement_matches_real_reference():
    d = np.load(GOLD)
    for g in range(6):
        adj = d[f"adj{g}"].astype(np.float64)
        labels = d[f"labels{g}"]
        C = int(labels.max()) + 1
        clusters = [np.nonzero(labels == k)[0].tolist() for k in range(C)]
        P, coarse = E.pooling_matrices(adj, clusters, 5)
        assert np.array_equal(coarse, d[f"coarse{g}"])
        for j in range(5):
            assert np.abs(P[j] - d[f"pool{g}"][j]).max() <= 1e-12, (g, j)
