"""N > 1 host logic on CPU (gloo, world_size 2): the all-gather-with-slice-backward used for the
global triplet loss and the flat-bucket gradient all-reduce.  The sharding is by graph with no other
exchange step, so data-parallel == single-process: the summed per-rank gradients of the GLOBAL loss
must equal the gradient a single process computes on the union of the shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _embed(w, x):              # stand-in for the per-rank packed forward (CPU: kernels need CUDA)
    return torch.tanh(x @ w)


def _global_loss(emb_all, trip, margin=1.5):
    a, p, n = emb_all[trip[:, 0]], emb_all[trip[:, 1]], emb_all[trip[:, 2]]
    dp = torch.sqrt(((a - p + 1e-6) ** 2).sum(-1)); dn = torch.sqrt(((a - n + 1e-6) ** 2).sum(-1))
    return torch.clamp((dp - dn) + margin, min=0).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "two-stage-gnn_b200"))
    from tsg.train import all_gather_rows, all_reduce_grads
    g = torch.Generator().manual_seed(0)
    M, Fin, D, T = 6, 5, 4, 9
    w = torch.randn(Fin, D, generator=g).requires_grad_(True)
    xs = torch.randn(world, M, Fin, generator=g)
    trip = torch.randint(0, world * M, (T, 3), generator=g)
    emb = _embed(w, xs[rank])
    emb_all = all_gather_rows(emb)
    assert emb_all.shape == (world * M, D)
    loss = _global_loss(emb_all, trip)
    loss.backward()
    all_reduce_grads([w])
    if rank == 0:
        torch.save(dict(grad=w.grad.clone(), loss=loss.detach(), emb_all=emb_all.detach()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_equals_single_process(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(0)
    M, Fin, D, T = 6, 5, 4, 9
    w = torch.randn(Fin, D, generator=g).requires_grad_(True)
    xs = torch.randn(world, M, Fin, generator=g)
    trip = torch.randint(0, world * M, (T, 3), generator=g)
    emb_all = _embed(w, xs.view(world * M, Fin))
    loss = _global_loss(emb_all, trip)
    loss.backward()
    assert torch.allclose(got["emb_all"], emb_all.detach(), atol=1e-7)
    assert torch.allclose(got["loss"], loss.detach(), atol=1e-7)
    assert torch.allclose(got["grad"], w.grad, atol=1e-6, rtol=1e-5)


def test_single_process_helpers_are_identity():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "two-stage-gnn_b200"))
    from tsg.train import all_gather_rows, all_reduce_grads
    e = torch.randn(3, 2, requires_grad=True)
    assert all_gather_rows(e) is e
    all_reduce_grads([e])          # no process group: no-op


def _worker_local(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "two-stage-gnn_b200"))
    from tsg.train import all_reduce_grads_and_loss
    g = torch.Generator().manual_seed(0)
    M, Fin, D = 6, 5, 4
    T = [4, 7]                                     # ranks hold DIFFERENT numbers of triplets
    w = torch.randn(Fin, D, generator=g).requires_grad_(True)
    b = torch.randn(D, generator=g).requires_grad_(True)
    xs = torch.randn(world, M, Fin, generator=g)
    trips = [torch.randint(0, M, (t, 3), generator=g) for t in T]
    emb = _embed(w, xs[rank]) + b
    loss = _global_loss(emb, trips[rank])          # mean over this rank's triplets only
    loss.backward()
    gl = all_reduce_grads_and_loss([w, b], loss.detach(), T[rank])
    if rank == 1:
        torch.save(dict(gw=w.grad.clone(), gb=b.grad.clone(), loss=gl.clone()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_one_collective_step_equals_single_process(tmp_path):
    """rank-local triplets: T_r-weighted all-reduce of [grads, loss, count] == loss / gradient of the union batch"""
    world, port, out = 2, _free_port(), str(tmp_path / "r1.pt")
    mp.spawn(_worker_local, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    g = torch.Generator().manual_seed(0)
    M, Fin, D = 6, 5, 4
    T = [4, 7]
    w = torch.randn(Fin, D, generator=g).requires_grad_(True)
    b = torch.randn(D, generator=g).requires_grad_(True)
    xs = torch.randn(world, M, Fin, generator=g)
    trips = [torch.randint(0, M, (t, 3), generator=g) for t in T]
    emb_all = _embed(w, xs.view(world * M, Fin)) + b
    trip_all = torch.cat([trips[r] + r * M for r in range(world)])
    loss = _global_loss(emb_all, trip_all)
    loss.backward()
    assert torch.allclose(got["loss"], loss.detach(), atol=1e-6)
    assert torch.allclose(got["gw"], w.grad, atol=1e-6, rtol=1e-5)
    assert torch.allclose(got["gb"], b.grad, atol=1e-6, rtol=1e-5)
