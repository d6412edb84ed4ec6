"""H1 native TU loader (host code in libtsg.so; no GPU needed) vs tests/golden/tu_ref.npz, the output of the REAL
reference `read_graphfile` on tests/golden/tu/TOY (oracle/make_golden_tu.py), and vs an independent numpy
restatement of PyG's TUDataset reader for the "pyg" mode."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PREFIX = os.path.join(HERE, "golden", "tu", "TOY", "TOY")


def _dense(c, g):
    n = c.num_nodes(g)
    e0, e1 = int(c.edge_ptr[g]), int(c.edge_ptr[g + 1])
    a = np.zeros((n, n), np.float32)
    a[c.row[e0:e1], c.col[e0:e1]] = 1.0
    return a


@pytest.mark.parametrize("tag,max_nodes", [("all", 0), ("max6", 6)])
def test_networkx_mode_matches_real_reference(tag, max_nodes):
    from tsg import tu
    ref = np.load(os.path.join(HERE, "golden", "tu_ref.npz"))
    corpus, attr, classes = tu.load(PREFIX, "networkx", max_nodes)
    assert corpus.num_graphs == int(ref[f"{tag}/num"])
    for g in range(corpus.num_graphs):
        adj = _dense(corpus, g)
        assert np.array_equal(adj, ref[f"{tag}/adj{g}"]), g           # node order + edge set + self loop
        assert int(corpus.y[g]) == int(ref[f"{tag}/y{g}"])
        n0, n1 = int(corpus.node_ptr[g]), int(corpus.node_ptr[g + 1])
        if n1 > n0:
            onehot = np.zeros((n1 - n0, corpus.num_node_labels), np.float32)
            onehot[np.arange(n1 - n0), corpus.node_label[n0:n1]] = 1.0
            assert np.array_equal(onehot, ref[f"{tag}/onehot{g}"])
            assert np.allclose(attr[n0:n1], ref[f"{tag}/feat{g}"], atol=1e-6)
    # edges are sorted by (row, col) inside every graph and symmetric
    for g in range(corpus.num_graphs):
        e0, e1 = int(corpus.edge_ptr[g]), int(corpus.edge_ptr[g + 1])
        code = corpus.row[e0:e1] * 10_000 + corpus.col[e0:e1]
        assert np.all(np.diff(code) > 0)
        assert np.array_equal(_dense(corpus, g), _dense(corpus, g).T)


def test_pyg_mode():
    """torch_geometric.io.read_tu_data semantics: all nodes kept in file order, self loops removed, coalesced,
    node label - min, graph label = rank among sorted distinct values."""
    from tsg import tu
    corpus, attr, classes = tu.load(PREFIX, "pyg")
    indic = np.loadtxt(PREFIX + "_graph_indicator.txt", dtype=np.int64)
    A = np.loadtxt(PREFIX + "_A.txt", dtype=np.int64, delimiter=",") - 1
    nl = np.loadtxt(PREFIX + "_node_labels.txt", dtype=np.int64)
    gl = np.loadtxt(PREFIX + "_graph_labels.txt", dtype=np.int64)
    G = gl.shape[0]
    assert corpus.num_graphs == G and classes == np.unique(gl).shape[0]
    assert np.array_equal(corpus.y, np.unique(gl, return_inverse=True)[1])
    assert np.array_equal(np.diff(corpus.node_ptr), np.bincount(indic - 1, minlength=G))
    assert np.array_equal(corpus.node_label, (nl - nl.min()).astype(np.int32))
    first = np.concatenate([[0], np.cumsum(np.bincount(indic - 1, minlength=G))])
    for g in range(G):
        m = (indic[A[:, 0]] - 1 == g) & (A[:, 0] != A[:, 1])
        e = np.unique(A[m] - first[g], axis=0)                       # coalesce = sort + dedup
        e0, e1 = int(corpus.edge_ptr[g]), int(corpus.edge_ptr[g + 1])
        assert np.array_equal(np.stack([corpus.row[e0:e1], corpus.col[e0:e1]], 1), e.reshape(-1, 2))
    assert attr.shape == (indic.shape[0], 3)


def test_missing_files_raise():
    from tsg import tu
    with pytest.raises(RuntimeError):
        tu.load(os.path.join(HERE, "golden", "tu", "NOPE", "NOPE"))


def test_loaded_corpus_feeds_the_compact_step_format():
    """H1 -> feeder: a corpus parsed from TU files becomes the compact host batch of run_from_host_compact; its edge
    lists are coalesced (sorted, loop free, symmetric), i.e. the form K1b's verified fast path takes."""
    import numpy as np
    from tsg import tu
    from tsg.feeder import compact_host_batch
    c, _, _ = tu.load(os.path.join(os.path.dirname(__file__), "golden", "tu", "TOY", "TOY"), "pyg")
    ids = np.array([3, 0, 3, 5], dtype=np.int64)
    b = compact_host_batch(c, ids, np.array([[0, 1, 3], [2, 3, 1]]), pin=False)
    n = np.diff(c.node_ptr)[ids]; e = np.diff(c.edge_ptr)[ids]
    assert np.array_equal(np.diff(b["node_ptr"]), n) and np.array_equal(np.diff(b["edge_ptr"]), e)
    assert b["label"].dtype.is_floating_point is False and b["label"].numel() == n.sum()
    assert b["row"].numel() == e.sum() and b["triplets"].shape == (2, 3)
    for k, g in enumerate(ids):
        sl = slice(int(b["edge_ptr"][k]), int(b["edge_ptr"][k + 1]))
        src = slice(int(c.edge_ptr[g]), int(c.edge_ptr[g + 1]))
        assert np.array_equal(b["row"][sl].numpy(), c.row[src]) and np.array_equal(b["col"][sl].numpy(), c.col[src])
        r, cc, nn = b["row"][sl].numpy().astype(np.int64), b["col"][sl].numpy().astype(np.int64), max(int(n[k]), 1)
        code = r * nn + cc
        assert np.all(np.diff(code) > 0) and np.all(r != cc)
        assert set((cc * nn + r).tolist()) == set(code.tolist())
    with pytest.raises(ValueError):
        compact_host_batch(c, ids, np.array([[0, 1, 4]]), pin=False)
