"""Seeded synthetic graph corpora of the dataset shapes the reference names
(`/root/reference/README.md:37-50`): DD 269 nodes / 676 edges, PROTEINS 39 / 73,
JAN. Y. 203 / 1866 (averages per graph, undirected edge counts).

Generator contract: sizes ~ clipped log-normal with the README mean; topology =
random spanning tree (connected => no isolated nodes, like `nx.from_edgelist` loading,
`Code/sage+gat+diffpool/load_data.py:89`) + uniformly random extra edges, no duplicates, no
self loops; stored symmetric (both directions) in lexicographic (row, col) order like
TUDataset's coalesced `edge_index`; node labels uniform over `num_node_labels` classes
(one-hot features, `Code/sag/train_triplet.py:161-163`); graph labels Bernoulli(0.5).

Pure numpy; no torch and no oracle imports (this is product-side input generation).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Sequence

import numpy as np

SHAPES: Dict[str, dict] = {
    #            mean nodes, mean undirected edges, clip range, #node labels
    "DD":       dict(n_mean=269.0, e_mean=676.0,  n_min=30, n_max=1000, num_node_labels=89),
    "PROTEINS": dict(n_mean=39.0,  e_mean=73.0,   n_min=4,  n_max=620,  num_node_labels=3),
    "JANY":     dict(n_mean=203.0, e_mean=1866.0, n_min=20, n_max=1000, num_node_labels=32),
}
SIGMA = 0.5   # log-normal shape parameter of the node-count distribution


@dataclass
class Corpus:
    """A set of graphs in concatenated (ragged) form; ids are LOCAL to each graph."""
    name: str
    node_ptr: np.ndarray      # int64 [G+1]
    edge_ptr: np.ndarray      # int64 [G+1]   directed edges
    row: np.ndarray           # int64 [sum E] local source id
    col: np.ndarray           # int64 [sum E] local target id
    node_label: np.ndarray    # int32 [sum n]
    y: np.ndarray             # int64 [G]
    num_node_labels: int

    @property
    def num_graphs(self) -> int:
        return int(self.y.shape[0])

    def num_nodes(self, g: int) -> int:
        return int(self.node_ptr[g + 1] - self.node_ptr[g])


def _one_graph(rng: np.random.Generator, n: int, m_undirected: int):
    """Random recursive spanning tree + extra edges. Returns (row, col) symmetric, lexicographic."""
    if n == 1:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    child = np.arange(1, n, dtype=np.int64)
    parent = (rng.random(n - 1) * child).astype(np.int64)          # parent < child
    lo, hi = parent, child
    max_m = n * (n - 1) // 2
    m = int(min(max(m_undirected, n - 1), max_m))
    extra = m - (n - 1)
    codes = lo * n + hi
    if extra > 0:
        have = set(codes.tolist())
        chosen = []
        while len(chosen) < extra:
            a = rng.integers(0, n, size=2 * (extra - len(chosen)) + 8)
            b = rng.integers(0, n, size=a.shape[0])
            for u, v in zip(a.tolist(), b.tolist()):
                if u == v:
                    continue
                c = (u * n + v) if u < v else (v * n + u)
                if c in have:
                    continue
                have.add(c); chosen.append(c)
                if len(chosen) == extra:
                    break
        codes = np.concatenate([codes, np.asarray(chosen, dtype=np.int64)])
    lo, hi = codes // n, codes % n
    r = np.concatenate([lo, hi]); c = np.concatenate([hi, lo])
    order = np.lexsort((c, r))
    return r[order], c[order]


def is_coalesced_symmetric(c: Corpus) -> bool:
    """True when every graph's edge list is in range, loop free, sorted by (row, col) without duplicates and symmetric --
    the TUDataset / TU-loader form `make_corpus` produces.  What `ops.CompactBatch.coalesced` promises (the kernels
    verify it again per graph on the device).  Vectorised: one lexicographic comparison and one sort over all edges."""
    E = int(c.edge_ptr[-1])
    if E == 0:
        return True
    g = np.repeat(np.arange(c.num_graphs, dtype=np.int64), np.diff(c.edge_ptr))
    n = np.diff(c.node_ptr)[g]
    row, col = c.row.astype(np.int64), c.col.astype(np.int64)
    if (row < 0).any() or (col < 0).any() or (row >= n).any() or (col >= n).any() or (row == col).any():
        return False
    nmax = int(n.max()) + 1
    key = (g * nmax + row) * nmax + col
    if not (np.diff(key) > 0).all():
        return False
    rkey = np.sort((g * nmax + col) * nmax + row)
    return bool(np.array_equal(key, rkey))


def make_corpus(shape: str, num_graphs: int, seed: int = 777) -> Corpus:
    """Graph g is drawn from `default_rng(seed + g)` so any subset is reproducible."""
    sp = SHAPES[shape]
    mu = np.log(sp["n_mean"]) - 0.5 * SIGMA * SIGMA
    ratio = sp["e_mean"] / sp["n_mean"]
    rows, cols, labels = [], [], []
    node_ptr = np.zeros(num_graphs + 1, np.int64)
    edge_ptr = np.zeros(num_graphs + 1, np.int64)
    y = np.zeros(num_graphs, np.int64)
    for g in range(num_graphs):
        rng = np.random.default_rng(seed + g)
        n = int(np.clip(np.rint(rng.lognormal(mu, SIGMA)), sp["n_min"], sp["n_max"]))
        m = int(np.rint(n * ratio))
        r, c = _one_graph(rng, n, m)
        rows.append(r); cols.append(c)
        labels.append(rng.integers(0, sp["num_node_labels"], size=n).astype(np.int32))
        y[g] = int(rng.random() < 0.5)
        node_ptr[g + 1] = node_ptr[g] + n
        edge_ptr[g + 1] = edge_ptr[g] + r.shape[0]
    return Corpus(shape, node_ptr, edge_ptr, np.concatenate(rows), np.concatenate(cols),
                  np.concatenate(labels), y, sp["num_node_labels"])


def tile_corpus(base: Corpus, num_graphs: int) -> Corpus:
    """Repeat a base corpus cyclically up to `num_graphs` graphs (used to reach the 1M-graph
    corpus of BASELINE config 5 without generating 1M distinct graphs on the host)."""
    reps = -(-num_graphs // base.num_graphs)
    ids = np.tile(np.arange(base.num_graphs), reps)[:num_graphs]
    return select(base, ids)


def select(c: Corpus, graph_ids: Sequence[int]) -> Corpus:
    ids = np.asarray(graph_ids, dtype=np.int64)
    n = c.node_ptr[ids + 1] - c.node_ptr[ids]
    e = c.edge_ptr[ids + 1] - c.edge_ptr[ids]
    node_ptr = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    edge_ptr = np.concatenate([[0], np.cumsum(e)]).astype(np.int64)
    nidx = _ragged_arange(c.node_ptr[ids], n)
    eidx = _ragged_arange(c.edge_ptr[ids], e)
    return Corpus(c.name, node_ptr, edge_ptr, c.row[eidx], c.col[eidx], c.node_label[nidx],
                  c.y[ids], c.num_node_labels)


def _ragged_arange(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    out_ptr = np.concatenate([[0], np.cumsum(lens)])[:-1]
    rep = np.repeat(starts - out_ptr, lens)
    return rep + np.arange(total, dtype=np.int64)


def pack(c: Corpus, graph_ids: Sequence[int] | None = None, one_hot: bool = True):
    """PyG `Batch.from_data_list` layout as numpy arrays:
    x [sum n, L] f32 one-hot (or int32 labels), edge_index [2, sum E] i64 (global ids),
    batch [sum n] i64, node_ptr [G+1] i64, y [G] i64."""
    if graph_ids is not None:
        c = select(c, graph_ids)
    G = c.num_graphs
    n = np.diff(c.node_ptr); e = np.diff(c.edge_ptr)
    off = np.repeat(c.node_ptr[:-1], e)
    edge_index = np.stack([c.row + off, c.col + off]).astype(np.int64)
    batch = np.repeat(np.arange(G, dtype=np.int64), n)
    if one_hot:
        x = np.zeros((int(c.node_ptr[-1]), c.num_node_labels), np.float32)
        x[np.arange(x.shape[0]), c.node_label] = 1.0
    else:
        x = c.node_label
    return dict(x=x, edge_index=edge_index, batch=batch, node_ptr=c.node_ptr.copy(), y=c.y.copy())


def sample_triplets(y: np.ndarray, num_triplets: int, seed: int = 0) -> np.ndarray:
    """Anchor i = graph (i mod G); positive = another graph of the same label, negative = a
    graph of the other label (the rule of `Code/sag/triplet_sampler.py:36-56`), drawn from a
    seeded generator instead of the reference's unseeded numpy/random state. int64 [T,3]."""
    rng = np.random.default_rng(seed)
    G = y.shape[0]
    by = {k: np.nonzero(y == k)[0] for k in (0, 1)}
    out = np.zeros((num_triplets, 3), np.int64)
    for t in range(num_triplets):
        a = t % G
        la = int(y[a])
        pos_pool, neg_pool = by[la], by[1 - la]
        p = a
        if pos_pool.shape[0] > 1:
            while p == a:
                p = int(pos_pool[rng.integers(0, pos_pool.shape[0])])
        nneg = int(neg_pool[rng.integers(0, neg_pool.shape[0])]) if neg_pool.shape[0] else a
        out[t] = (a, p, nneg)
    return out
