"""layers.py:2 `from torch_geometric.nn.pool.topk_pool import topk, filter_adj` -> oracle.pyg_ref."""
from oracle import pyg_ref as R


def topk(x, ratio, batch, min_score=None, tol=1e-7):
    assert min_score is None
    return R.topk(x, ratio, batch)


def filter_adj(edge_index, edge_attr, perm, num_nodes=None):
    return R.filter_adj(edge_index, edge_attr, perm, num_nodes)
