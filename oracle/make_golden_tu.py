"""Generates tests/golden/tu/ (a small TU-format dataset) and tests/golden/tu_ref.npz = what the REAL reference
loader `/root/reference/Code/sage+gat+diffpool/load_data.py::read_graphfile` returns for it (networkx graphs:
node order, one-hot labels, graph labels, adjacency), so the native H1 loader can be checked against it.
networkx 3.x: the reference's `float(nx.__version__)` needs a two-component version string -> patched.
Run from the repo root:  python oracle/make_golden_tu.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/Code/sage+gat+diffpool"
NAME = "TOY"


def write_dataset(d):
    rng = np.random.default_rng(5)
    os.makedirs(os.path.join(d, NAME), exist_ok=True)
    pre = os.path.join(d, NAME, NAME)
    sizes = [5, 1, 7, 4, 9, 3]                   # graph 2 is a single node without edges
    glabels = [3, 3, -1, 7, -1, 3]               # non-consecutive labels, first-appearance order != sorted order
    indic, nlab, edges = [], [], []
    base = 0
    for g, n in enumerate(sizes):
        indic += [g + 1] * n
        nlab += [int(v) for v in rng.integers(1, 5, size=n)]
        und = set()
        order = rng.permutation(n)               # a tree in random order so that first-appearance order != id order
        for k in range(1, n):
            a, b = int(order[k]), int(order[rng.integers(0, k)])
            und.add((min(a, b), max(a, b)))
        for _ in range(n // 2):
            a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
            if a != b:
                und.add((min(a, b), max(a, b)))
        lst = [(a, b) for a, b in und] + [(b, a) for a, b in und]
        lst = [lst[i] for i in rng.permutation(len(lst))]          # TU files are sorted, but nothing relies on it
        if g == 4:
            lst.append((2, 2))                                       # a self loop
            lst.append(lst[0])                                       # a duplicate line
        if g == 3 and n > 3:                                         # node 3 of graph 4 stays isolated
            lst = [(a, b) for a, b in lst if a != 3 and b != 3]
        edges += [(base + a + 1, base + b + 1) for a, b in lst]
        base += n
    open(pre + "_graph_indicator.txt", "w").write("".join(f"{v}\n" for v in indic))
    open(pre + "_graph_labels.txt", "w").write("".join(f"{v}\n" for v in glabels))
    open(pre + "_node_labels.txt", "w").write("".join(f"{v}\n" for v in nlab))
    open(pre + "_A.txt", "w").write("".join(f"{a}, {b}\n" for a, b in edges))
    attrs = rng.normal(size=(base, 3)).astype(np.float32)
    open(pre + "_node_attributes.txt", "w").write("".join(", ".join(f"{x:.6f}" for x in r) + "\n" for r in attrs))


def main():
    gold = os.path.join(ROOT, "tests", "golden", "tu")
    write_dataset(gold)
    import networkx as nx
    nx.__version__ = "3.6"
    sys.path.insert(0, REF)
    import load_data
    out = {}
    for tag, mx in (("all", None), ("max6", 6)):
        graphs = load_data.read_graphfile(gold, NAME, max_nodes=mx)
        out[f"{tag}/num"] = np.asarray(len(graphs))
        for i, G in enumerate(graphs):
            nodes = list(G.nodes())
            assert nodes == list(range(len(nodes)))
            out[f"{tag}/adj{i}"] = np.asarray(nx.to_numpy_array(G, nodelist=nodes), np.float32) if nodes else np.zeros((0, 0), np.float32)
            out[f"{tag}/y{i}"] = np.asarray(G.graph["label"])
            out[f"{tag}/onehot{i}"] = (np.asarray([G.nodes[u]["label"] for u in nodes], np.float32) if nodes
                                       else np.zeros((0, 0), np.float32))
            out[f"{tag}/feat{i}"] = (np.asarray([G.nodes[u]["feat"] for u in nodes], np.float32) if nodes
                                     else np.zeros((0, 0), np.float32))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tu_ref.npz"), **out)
    print("wrote tests/golden/tu_ref.npz:", int(out["all/num"]), "graphs;", int(out["max6/num"]), "with max_nodes=6")


if __name__ == "__main__":
    main()
