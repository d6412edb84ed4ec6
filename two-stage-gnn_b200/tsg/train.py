"""2stg training step over packed batches: the public call a user of this framework makes.

Mirrors the loop of Code/sag/train_triplet.py:198-214 (TNet forward -> MarginRankingLoss ->
backward -> Adam step) with two differences that are the point of this framework: the step runs
ONE packed forward over all 3T graphs of T triplets instead of 3 single-graph forwards per
triplet, and it is data-parallel: every rank embeds its own graphs and evaluates its own triplets; ONE
all-reduce (NCCL) per step of the flat bucket [T_r * gradients, T_r * loss, T_r] yields the global-batch mean
loss and its gradient on every rank.  Graphs are independent, so there is no other exchange step
(`all_gather_rows` stays for triplets that reference graphs of other ranks).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import ops


class _AllGatherRows(torch.autograd.Function):
    """emb_local [M, D] -> emb_global [world*M, D].  Every rank evaluates the SAME global loss on
    the gathered matrix, so d(loss)/d(emb_global) is complete on every rank and the backward is a
    slice (no reduce-scatter needed)."""

    @staticmethod
    def forward(ctx, emb, group):
        world = dist.get_world_size(group)
        ctx.rank, ctx.m = dist.get_rank(group), emb.size(0)
        out = torch.empty(world * emb.size(0), emb.size(1), dtype=emb.dtype, device=emb.device)
        dist.all_gather_into_tensor(out, emb.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank * ctx.m:(ctx.rank + 1) * ctx.m].contiguous(), None


def all_gather_rows(emb: torch.Tensor, group=None) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return emb
    return _AllGatherRows.apply(emb, group)


def all_reduce_grads(params, group=None) -> None:
    """One flat-bucket all-reduce(SUM) of every parameter gradient (33-361 KB: latency bound)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g)); off += n
    torch._foreach_copy_(grads, views)          # one multi-tensor kernel instead of one copy per parameter


def all_reduce_grads_and_loss(params, loss_local: torch.Tensor, num_local: int, group=None) -> torch.Tensor:
    """The step's ONE collective when every rank's triplets index its own graphs.  `loss_local` is the MEAN hinge over
    this rank's `num_local` triplets and the parameters hold its gradient; the global batch's mean loss is
    sum_r(T_r * loss_r) / sum_r(T_r) and its gradient the same combination of the per-rank gradients, so one
    all-reduce(SUM) of [T_r * grads..., T_r * loss_r, T_r] carries everything -- no embedding all-gather, no triplet
    all-gather, nothing to wait for before the backward pass.  Returns the global mean loss (on the device)."""
    grads = [p.grad for p in params if p.grad is not None]
    tail = torch.stack([loss_local.detach().reshape(()), torch.ones((), dtype=loss_local.dtype, device=loss_local.device)])
    flat = torch.cat([g.reshape(-1) for g in grads] + [tail])
    flat.mul_(float(num_local))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(flat[-1].clone())
    views, off = [], 0
    for g in grads:
        n = g.numel()
        views.append(flat[off:off + n].view_as(g)); off += n
    if grads:
        torch._foreach_copy_(grads, views)
    return flat[-2]


class CapturedStep:
    """A whole training step (forward through the tsg operators, loss, backward, clipping, optimiser) captured ONCE into a
    CUDA graph and replayed -- for steps whose shapes do not change between iterations (the reference's "original"
    setting trains on the same packed corpus every epoch, Code/sage+gat+diffpool/train.py:85-152): ~200 launches of a few
    microseconds each become one graph launch.  This is what the C ABI's contract buys (include/tsg.h: every compute
    entry point only enqueues on the caller's stream, never allocates, never synchronises; tsg_init_device() runs
    before the capture): the library's kernels, memsets and last-CTA reduction tickets are all capturable.

    `body()` must read its inputs from tensors that stay alive and in place (update them with copy_()), use an optimiser
    built with `capturable=True`, and return a tensor (the loss); its value after each replay() is in `.out`."""

    def __init__(self, body, warmup: int = 3):
        from . import _lib
        _lib.init_device()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                    # warm-up on a side stream (allocator, lazy inits, autotuned paths)
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = body()

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


def _check_status():
    from . import nn as _nn
    _nn.check_fused_status()


class TripletTrainer:
    """model: tsg.nn.PackedSAGNet (or any module with the same forward signature)."""

    def __init__(self, model: torch.nn.Module, lr: float = 5e-4, weight_decay: float = 1e-4,
                 margin: float = 1.5, group=None):
        self.model, self.margin, self.group = model, margin, group
        # Code/sag/train_triplet.py:191
        params = list(model.parameters())
        # one fused multi-tensor kernel per step when the parameters live on the GPU (same update rule as the
        # reference's torch.optim.Adam; the foreach path costs ~8 launches and 0.15 ms of host time per step)
        fused = bool(params) and all(p.is_cuda for p in params)
        self.opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, fused=fused)

    def step(self, x: torch.Tensor, edge_index: torch.Tensor, node_ptr_host: np.ndarray,
             triplets: torch.Tensor) -> torch.Tensor:
        """x/edge_index/triplets on the device.  `triplets` [T,3] index rows of THIS rank's embedding matrix; with
        world > 1 the returned loss is the mean over the GLOBAL batch and the parameters receive its gradient."""
        self.model.train()
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if (edge_index is None and getattr(self.model, "native_step_supported", None)
                and self.model.native_step_supported(x, node_ptr_host)):
            # K14: forward, loss and backward are ONE C-ABI call; gradients land in views of one flat buffer
            loss = self.model.native_step(x, node_ptr_host, triplets, self.margin)
            if world > 1:
                flat, _ = self.model._flat_grads()
                flat[-1:].fill_(1.0)           # (NOT flat[-1] = 1.0: a Python scalar store is a blocking H2D copy -- measured
                                               # 1.2 ms of host stall per step, 1.43 -> 1.87 ms at any N > 1: profiles/tools/scale_probe2.py)
                flat.mul_(float(triplets.size(0)))                       # [T_r * grads, T_r * loss, T_r]
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(flat[-1].clone())
            loss = loss.clone()                                           # the buffer is rewritten by the next step
            self.opt.step()
            return loss
        emb = self.model(x, edge_index, node_ptr_host)
        loss, _, _ = ops.triplet_loss(emb, triplets, self.margin)         # mean over THIS rank's triplets
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:
            # triplets are rank-local by contract, so the global-batch loss and gradient are the T_r-weighted means of
            # the per-rank ones: one all-reduce per step (the all-gather formulation, kept as all_gather_rows for
            # triplets that cross ranks, cost two more collectives and a synchronisation before the backward pass)
            loss = all_reduce_grads_and_loss(list(self.model.parameters()), loss, int(triplets.size(0)), self.group)
        self.opt.step()
        return loss.detach()

    def step_allgather(self, x, edge_index, node_ptr_host: np.ndarray, triplets_global: torch.Tensor) -> torch.Tensor:
        """The north-star's formulation for triplets that CROSS ranks (the reference sampler draws positives and negatives
        from the whole training set, Code/sag/triplet_sampler.py:36-56): every rank embeds its own M graphs, the
        embeddings are all-gathered into [world * M, D] (row r * M + i = graph i of rank r), every rank evaluates the SAME
        global loss over `triplets_global` [T_g, 3] (indices into the gathered matrix, identical on all ranks), backward
        takes this rank's slice of d(loss)/d(embeddings), and one all-reduce(SUM) completes the parameter gradient."""
        self.model.train()
        if (edge_index is None and getattr(self.model, "native_step_supported", None)
                and self.model.native_step_supported(x, node_ptr_host)):
            # native halves (K14): forward to the embeddings, gather, loss + its gradient on the gathered matrix (K9),
            # this rank's slice back into the native backward
            emb, ctx = self.model.native_forward(x, node_ptr_host)
            world = dist.get_world_size(self.group) if dist.is_initialized() else 1
            if world > 1:
                emb_g = torch.empty(world * emb.size(0), emb.size(1), dtype=emb.dtype, device=emb.device)
                dist.all_gather_into_tensor(emb_g, emb, group=self.group)
            else:
                emb_g = emb.clone()
            emb_g.requires_grad_(True)
            loss, _, _ = ops.triplet_loss(emb_g, triplets_global, self.margin)
            loss.backward()
            r = dist.get_rank(self.group) if world > 1 else 0
            self.model.native_backward(ctx, emb_g.grad[r * emb.size(0):(r + 1) * emb.size(0)])
            if world > 1:
                flat, _ = self.model._flat_grads()
                dist.all_reduce(flat[:-2], op=dist.ReduceOp.SUM, group=self.group)
            self.opt.step()
            return loss.detach()
        emb = self.model(x, edge_index, node_ptr_host)
        emb_g = all_gather_rows(emb, self.group)
        loss, _, _ = ops.triplet_loss(emb_g, triplets_global, self.margin)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        all_reduce_grads(list(self.model.parameters()), self.group)
        self.opt.step()
        return loss.detach()

    def step_from_host(self, x_host: torch.Tensor, edge_index_host: torch.Tensor,
                       node_ptr_host: np.ndarray, triplets_host: torch.Tensor, device) -> float:
        """End-to-end call with (pinned) HOST buffers: H2D copies, the step, and the loss read back."""
        x = x_host.to(device, non_blocking=True)
        ei = edge_index_host.to(device, non_blocking=True)
        tr = triplets_host.to(device, non_blocking=True)
        return float(self.step(x, ei, node_ptr_host, tr).item())

    def run_from_host(self, host_batches, device) -> list:
        """Pipelined end-to-end loop over an iterable of HOST batches (dicts with pinned `x`, `edge_index`,
        `triplets` and numpy `node_ptr`): the H2D copies of batch i+1 run on a copy stream while the step
        on batch i computes, and every step's loss is read back into pinned host memory without stalling
        the enqueue thread.  Same arithmetic as step_from_host; returns the per-step losses."""
        cur = torch.cuda.current_stream(device)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(device)

        def upload(b):
            copy.wait_stream(cur)      # staging memory freed by earlier steps must be done being read
            with torch.cuda.stream(copy):
                x = b["x"].to(device, non_blocking=True)
                ei = b["edge_index"].to(device, non_blocking=True)
                tr = b["triplets"].to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return x, ei, b["node_ptr"], tr, ev

        it = iter(host_batches)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return []
        host_losses = []
        while nxt is not None:
            x, ei, nptr, tr, ev = nxt
            try:
                nxt = upload(next(it))
            except StopIteration:
                nxt = None
            cur.wait_event(ev)
            for t in (x, ei, tr):
                t.record_stream(cur)
            loss = self.step(x, ei, nptr, tr)
            h = torch.empty((), dtype=torch.float32, pin_memory=True)
            h.copy_(loss, non_blocking=True)
            host_losses.append(h)
        cur.synchronize()
        _check_status()
        return [float(h) for h in host_losses]

    def run_from_host_compact(self, host_batches, device, num_node_labels: int, expand: bool = False) -> list:
        """Pipelined end-to-end loop over COMPACT host batches: what a TU dataset actually stores for a graph --
        node labels and the edge list -- instead of the fp32 one-hot matrix and int64 edge_index the PyG surface
        materialises.  Each batch: dict(label int32 [N], row / col int32 [E] LOCAL node ids, node_ptr / edge_ptr
        int64 numpy [G+1], triplets int64 [T,3]), tensors pinned.  Per step the H2D is 4N + 8E + 16G bytes
        (43 MB instead of 423 MB for the bench batch).  The batch goes to the model as an `ops.CompactBatch`:
        PackedSAGNet's executor consumes labels and local endpoints directly (K3c gather / segment sum for conv1,
        K1b on int32 local ids), so neither the fp32 one-hot x nor the int64 edge_index is ever written.
        `expand=True` restores the earlier behaviour: `tsg_pack_batch` (K0) materialises both on the GPU first
        (models without a compact path do that themselves through `CompactBatch.expand`)."""
        from ._lib import call, ptr, stream_ptr
        cur = torch.cuda.current_stream(device)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(device)

        def upload(b):
            copy.wait_stream(cur)
            G = b["node_ptr"].shape[0] - 1
            meta = b.get("_meta")
            if meta is None:          # pinned once per host batch, reused when the batch comes round again
                meta = b["_meta"] = torch.from_numpy(np.concatenate([b["node_ptr"], b["edge_ptr"]])).pin_memory()
            with torch.cuda.stream(copy):
                t = {k: b[k].to(device, non_blocking=True) for k in ("label", "row", "col", "triplets")}
                t["meta"] = meta.to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            return t, b["node_ptr"], b["edge_ptr"], ev, b.get("coalesced", False)

        it = iter(host_batches)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return []
        host_losses = []
        while nxt is not None:
            t, nptr, eptr, ev, b_coalesced = nxt
            try:
                nxt = upload(next(it))
            except StopIteration:
                nxt = None
            cur.wait_event(ev)
            for v in t.values():
                v.record_stream(cur)
            G, N, E = nptr.shape[0] - 1, int(nptr[-1]), int(eptr[-1])
            d_nptr, d_eptr = t["meta"][:G + 1], t["meta"][G + 1:]
            cb = ops.CompactBatch(t["label"], t["row"], t["col"], d_nptr, d_eptr, num_node_labels,
                                  int(np.diff(eptr).max()) if G else 0, bool(b_coalesced))
            if expand or not getattr(self.model, "accepts_compact", False):
                x, ei = cb.expand()
            else:
                x, ei = cb, None
            loss = self.step(x, ei, nptr, t["triplets"])
            h = torch.empty((), dtype=torch.float32, pin_memory=True)
            h.copy_(loss, non_blocking=True)
            host_losses.append(h)
        cur.synchronize()
        _check_status()
        return [float(h) for h in host_losses]

    def run_from_ids(self, corpus, id_batches) -> list:
        """Pipelined end-to-end loop against an HBM-resident corpus (tsg.feeder.DeviceCorpus): every step the host
        provides only what a sampler produces -- the step's graph ids and its triplet index rows -- as
        (graph_ids int64 numpy [B], triplets int64 pinned tensor [T, 3]).  Per step, INSIDE the loop: the packed
        offsets of the chosen graphs are computed on the host (two cumsums over B sizes), ids + offsets + triplets go
        up on the copy stream, the batch is assembled on the GPU from the resident corpus (K0 compact gather), the
        training step runs, and the loss is read back into pinned memory without stalling the enqueue thread.
        Nothing is pre-packed or cached across steps.  Returns the per-step losses."""
        dev = corpus.device
        cur = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
        compact = corpus.label is not None and getattr(self.model, "accepts_compact", False)

        def stage(item):
            ids_host, trip_host = item
            ids, nptr, eptr = corpus.offsets(ids_host)
            need = ids.shape[0] + nptr.shape[0] + eptr.shape[0]
            slot = self._ring_pos = (getattr(self, "_ring_pos", -1) + 1) % 4
            ring = self.__dict__.setdefault("_meta_ring", [None] * 4)
            if ring[slot] is None or ring[slot][0].numel() < need:       # pinned staging ring: allocated once, reused
                ring[slot] = [torch.empty(max(need, 1024), dtype=torch.int64, pin_memory=True), None]
            buf, busy = ring[slot]
            if busy is not None:
                busy.synchronize()             # the upload issued from this slot four stages ago has left it
            meta_h = buf[:need]
            np.concatenate([ids, nptr, eptr], out=meta_h.numpy())
            copy.wait_stream(cur)
            with torch.cuda.stream(copy):
                meta = meta_h.to(dev, non_blocking=True)
                tr = trip_host.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            ring[slot][1] = ev
            return meta, tr, ids.shape[0], nptr, eptr, ev, meta_h

        it = iter(id_batches)
        try:
            nxt = stage(next(it))
        except StopIteration:
            return []
        host_losses = []
        while nxt is not None:
            meta, tr, B, nptr, eptr, ev, keep = nxt
            try:
                nxt = stage(next(it))
            except StopIteration:
                nxt = None
            cur.wait_event(ev)
            meta.record_stream(cur); tr.record_stream(cur)
            if compact:
                (x, _), ei = corpus.pack_compact_staged(meta, B, nptr, eptr), None
            else:
                x, ei, _ = corpus.pack_staged(meta, B, nptr, eptr)
            loss = self.step(x, ei, nptr, tr)
            h = torch.empty((), dtype=torch.float32, pin_memory=True)
            h.copy_(loss, non_blocking=True)
            host_losses.append(h)
        cur.synchronize()
        _check_status()
        return [float(h) for h in host_losses]

    def step_from_ids(self, corpus, graph_ids_host: np.ndarray, triplets_host: torch.Tensor) -> float:
        """End-to-end call against an HBM-resident corpus (tsg.feeder.DeviceCorpus): the host sends the
        step's graph ids + triplet index list, the batch is assembled on the GPU, the loss is read back."""
        if corpus.label is not None and getattr(self.model, "accepts_compact", False):
            (x, nptr), ei = corpus.pack_compact(graph_ids_host), None
        else:
            x, ei, nptr = corpus.pack(graph_ids_host)
        tr = triplets_host.to(corpus.device, non_blocking=True)
        loss = float(self.step(x, ei, nptr, tr).item())
        _check_status()
        return loss
