"""TEST INFRASTRUCTURE ONLY -- a `torch_geometric` package whose operators are the CPU oracle
(oracle/pyg_ref.py).  It exists so that the UNMODIFIED reference glue (Code/sag/layers.py, network.py)
can be executed on CPU in this container: oracle/make_golden_sag.py imports the real `network.Net` on top
of it and stores what it computes (tests/golden/sag_glue.npz).  That pins the GLUE above the operators
(which conv feeds which pool, squeeze/view, the raw score ranking, readout order, the x1+x2+x3 sum, the
head) to the reference's own code; the operators underneath remain the restatement of PyG 1.6.3
(PARITY UNPINNED against upstream PyG binaries, see oracle/pyg_ref.py).  Never imported by the product."""
__version__ = "1.6.3-oracle"
