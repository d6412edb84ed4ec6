"""torch_geometric.nn.pool.topk_pool.{topk, filter_adj} as called from Code/sag/layers.py:20,23."""
import torch

from tsg import ops


def topk(x, ratio, batch, min_score=None, tol=1e-7):
    if min_score is not None:
        raise NotImplementedError("min_score is not used by the reference (layers.py:20)")
    x = x.view(-1)
    G = int(batch.max().item()) + 1 if batch.numel() else 0          # upstream syncs here too
    gptr = ops.batch_to_ptr(batch, G)
    kptr = ops.topk_sizes(gptr, ratio)
    return ops.topk(x, gptr, kptr, int(kptr[-1].item()))


def filter_adj(edge_index, edge_attr, perm, num_nodes=None):
    if edge_attr is not None:
        raise NotImplementedError("edge_attr is None everywhere in the reference (layers.py:23)")
    if num_nodes is None:
        num_nodes = int(edge_index.max().item()) + 1
    el, _ = ops.filter_adj(ops.EdgeList.from_edge_index(edge_index), perm, int(num_nodes))
    return el.edge_index(), None
