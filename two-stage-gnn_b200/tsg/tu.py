"""TU-format dataset loader (H1): text files -> `tsg.synth.Corpus` through the native one-pass
parser in libtsg.so.  mode "networkx" = the dense directories' `read_graphfile` semantics
(Code/sage+gat+diffpool/load_data.py:12-126), mode "pyg" = TUDataset's (Code/sag/train*.py:161-163)."""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .synth import Corpus

MODES = {"networkx": 0, "pyg": 1}


def find_prefix(root: str, name: str) -> Optional[str]:
    """<root>/<name>/<name>, <root>/<name>/raw/<name> (PyG layout) or <root>/<name>: first that has _A.txt."""
    for cand in (os.path.join(root, name, name), os.path.join(root, name, "raw", name), os.path.join(root, name)):
        if os.path.exists(cand + "_A.txt"):
            return cand
    return None


def load(prefix: str, mode: str = "pyg", max_nodes: int = 0) -> Tuple[Corpus, Optional[np.ndarray], int]:
    """Returns (corpus, node attributes [N, d] or None, number of graph classes)."""
    h = ctypes.c_void_p()
    rc = _lib.lib.tsg_tu_load(prefix.encode(), MODES[mode], int(max_nodes), ctypes.byref(h))
    if rc != 0:
        raise RuntimeError(f"tsg_tu_load failed ({rc}): {_lib.last_error()}")
    try:
        sizes = (ctypes.c_int64 * 6)()
        _lib.lib.tsg_tu_sizes(h, sizes)
        G, N, E, L, D, C = [int(v) for v in sizes]
        node_ptr = np.zeros(G + 1, np.int64); edge_ptr = np.zeros(G + 1, np.int64)
        row = np.zeros(E, np.int64); col = np.zeros(E, np.int64)
        label = np.zeros(N, np.int32); y = np.zeros(G, np.int64)
        attr = np.zeros((N, D), np.float32) if D > 0 else None
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None and a.size else None
        rc = _lib.lib.tsg_tu_fill(h, p(node_ptr), p(edge_ptr), p(row), p(col), p(label), p(y), p(attr))
        if rc != 0:
            raise RuntimeError(f"tsg_tu_fill failed ({rc}): {_lib.last_error()}")
    finally:
        _lib.lib.tsg_tu_free(h)
    name = os.path.basename(prefix)
    return Corpus(name, node_ptr, edge_ptr, row, col, label, y, max(L, 1)), attr, C


def write(corpus: Corpus, root: str, name: str) -> str:
    """Export a Corpus in the TU text format (<root>/<name>/<name>_{A,graph_indicator,graph_labels,node_labels}.txt,
    1-based global node ids, one directed edge per line): the inverse of `load`, used to hand synthetic corpora to the
    UNMODIFIED reference scripts through their own loaders (tests/test_launcher.py).  Returns the file prefix."""
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    prefix = os.path.join(d, name)
    off = np.repeat(corpus.node_ptr[:-1], np.diff(corpus.edge_ptr))
    np.savetxt(prefix + "_A.txt", np.stack([corpus.row + off + 1, corpus.col + off + 1], 1), fmt="%d", delimiter=", ")
    np.savetxt(prefix + "_graph_indicator.txt", np.repeat(np.arange(1, corpus.num_graphs + 1), np.diff(corpus.node_ptr)), fmt="%d")
    np.savetxt(prefix + "_graph_labels.txt", corpus.y, fmt="%d")
    np.savetxt(prefix + "_node_labels.txt", corpus.node_label, fmt="%d")
    return prefix
