"""Device-resident corpus + GPU batch assembly.

The reference moves every graph of every triplet host->device each step (and, in the dense
directories, rebuilds 1000x1000 tensors from numpy: 115-125 ms per graph).  Here the corpus is uploaded
once (DD: 1,168 graphs = 14 MB) and a step ships only graph ids; `tsg_pack_batch` writes the packed
batch at HBM speed."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call, ptr, stream_ptr
from .synth import Corpus


class _PinnedRing:
    """Four reusable page-locked int64 staging buffers: a per-step `.pin_memory()` on a fresh tensor is a host
    allocation + registration every call.  A slot is reused only after the upload issued from it has left it."""

    def __init__(self):
        self.slots, self.pos = [None] * 4, -1

    def stage(self, parts, device) -> torch.Tensor:
        need = sum(int(p.shape[0]) for p in parts)
        self.pos = (self.pos + 1) % 4
        slot = self.slots[self.pos]
        if slot is None or slot[0].numel() < need:
            slot = self.slots[self.pos] = [torch.empty(max(need, 1024), dtype=torch.int64, pin_memory=True), None]
        buf, busy = slot
        if busy is not None:
            busy.synchronize()
        host = buf[:need]
        np.concatenate(parts, out=host.numpy())
        dev = host.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot[1] = ev
        return dev


class DeviceCorpus:
    def __init__(self, corpus: Corpus, device, dense_x: np.ndarray | None = None):
        self.device = device
        self.n = np.diff(corpus.node_ptr).astype(np.int64)          # host copies: sizes drive the offsets
        self.e = np.diff(corpus.edge_ptr).astype(np.int64)
        self.num_graphs = corpus.num_graphs
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(device)
        self.node_ptr, self.edge_ptr = t(corpus.node_ptr, np.int64), t(corpus.edge_ptr, np.int64)
        self.row, self.col = t(corpus.row, np.int32), t(corpus.col, np.int32)
        from .synth import is_coalesced_symmetric
        self.coalesced = is_coalesced_symmetric(corpus)      # checked once, here; the kernels re-verify per graph
        if dense_x is None:
            self.label, self.x, self.feat = t(corpus.node_label, np.int32), None, corpus.num_node_labels
        else:
            self.label, self.x, self.feat = None, t(dense_x, np.float32), int(dense_x.shape[1])

    @classmethod
    def from_device(cls, label: torch.Tensor, row: torch.Tensor, col: torch.Tensor, node_ptr_host: np.ndarray,
                    edge_ptr_host: np.ndarray, num_labels: int, coalesced: bool) -> "DeviceCorpus":
        """A corpus whose arrays already live in HBM (e.g. a shard assembled on the GPU from a smaller base corpus:
        bench.py's 1 M-graph config-5 corpus); only the two offset arrays come from the host."""
        self = cls.__new__(cls)
        self.device = label.device
        self.n = np.diff(node_ptr_host).astype(np.int64); self.e = np.diff(edge_ptr_host).astype(np.int64)
        self.num_graphs = int(self.n.shape[0])
        self.node_ptr = torch.from_numpy(np.ascontiguousarray(node_ptr_host, dtype=np.int64)).to(self.device)
        self.edge_ptr = torch.from_numpy(np.ascontiguousarray(edge_ptr_host, dtype=np.int64)).to(self.device)
        self.row, self.col, self.label, self.x, self.feat = row, col, label, None, int(num_labels)
        self.coalesced = bool(coalesced)
        return self

    def offsets(self, ids_host: np.ndarray):
        """Packed offsets for a list of graph ids (host numpy; int64 [B+1] each)."""
        ids = np.asarray(ids_host, dtype=np.int64)
        nptr = np.zeros(ids.shape[0] + 1, np.int64); np.cumsum(self.n[ids], out=nptr[1:])
        eptr = np.zeros(ids.shape[0] + 1, np.int64); np.cumsum(self.e[ids], out=eptr[1:])
        return ids, nptr, eptr

    def _stage(self, ids_host: np.ndarray):
        """ids + packed offsets of the chosen graphs as ONE small pinned upload: (meta on the device, B, nptr, eptr)."""
        ids, nptr, eptr = self.offsets(ids_host)
        ring = self.__dict__.get("_ring")
        if ring is None:
            ring = self.__dict__["_ring"] = _PinnedRing()
        return ring.stage([ids, nptr, eptr], self.device), ids.shape[0], nptr, eptr

    def pack(self, ids_host: np.ndarray):
        """-> (x [sum n, F] f32, edge_index [2, sum E] i64, node_ptr_host).  One small H2D (ids + offsets)."""
        return self.pack_staged(*self._stage(ids_host))

    def pack_staged(self, meta: torch.Tensor, B: int, nptr: np.ndarray, eptr: np.ndarray):
        """`pack` when ids + offsets are already on the device (`meta` = [ids | nptr | eptr], int64)."""
        d_ids, d_nptr, d_eptr = meta[:B], meta[B:2 * B + 1], meta[2 * B + 1:]
        N, E = int(nptr[-1]), int(eptr[-1])
        x = torch.empty(N, self.feat, dtype=torch.float32, device=self.device)
        ei = torch.empty(2, E, dtype=torch.int64, device=self.device)
        call("tsg_pack_batch", ptr(d_ids), ptr(d_nptr), ptr(d_eptr), B, ptr(self.node_ptr), ptr(self.edge_ptr),
             ptr(self.row), ptr(self.col), ptr(self.label), ptr(self.x), self.feat, ptr(x), ptr(ei[0]), ptr(ei[1]),
             stream_ptr())
        return x, ei, nptr

    def pack_compact(self, ids_host: np.ndarray):
        """-> (ops.CompactBatch, node_ptr_host): labels + graph-local int32 endpoints of the chosen graphs, gathered
        on the GPU (4N + 8E bytes written; PackedSAGNet consumes the batch without expanding it)."""
        return self.pack_compact_staged(*self._stage(ids_host))

    def pack_compact_staged(self, meta: torch.Tensor, B: int, nptr: np.ndarray, eptr: np.ndarray):
        from .ops import CompactBatch
        if self.label is None:
            raise RuntimeError("tsg: the compact batch form needs categorical node labels (corpus holds dense features)")
        d_ids, d_nptr, d_eptr = meta[:B], meta[B:2 * B + 1], meta[2 * B + 1:]
        N, E = int(nptr[-1]), int(eptr[-1])
        label = torch.empty(N, dtype=torch.int32, device=self.device)
        rc = torch.empty(2, E, dtype=torch.int32, device=self.device)
        call("tsg_pack_batch_compact", ptr(d_ids), ptr(d_nptr), ptr(d_eptr), B, ptr(self.node_ptr), ptr(self.edge_ptr),
             ptr(self.row), ptr(self.col), ptr(self.label), ptr(label), ptr(rc[0]), ptr(rc[1]), stream_ptr())
        return CompactBatch(label, rc[0], rc[1], d_nptr, d_eptr, self.feat, int(np.diff(eptr).max()) if B else 0,
                            self.coalesced), nptr


class DeviceRagged:
    """Per-graph variable-length rows of 32-bit values resident in HBM (cluster labels per node, pooling weights per
    cluster, ...), gathered by graph id with the same kernel as the corpus itself (K0 compact with no edges)."""

    def __init__(self, ptr_host: np.ndarray, values: torch.Tensor):
        assert values.dtype in (torch.int32, torch.float32)
        self.device = values.device
        self.len = np.diff(ptr_host).astype(np.int64)
        self.ptr = torch.from_numpy(np.ascontiguousarray(ptr_host, dtype=np.int64)).to(self.device)
        self.values = values.contiguous()
        self._zero_ptr = torch.zeros(self.len.shape[0] + 1, dtype=torch.int64, device=self.device)
        self._dummy = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._ring = _PinnedRing()

    def gather(self, ids_host: np.ndarray):
        """-> (values of the chosen graphs, concatenated; their offsets on the host [B+1])."""
        ids = np.asarray(ids_host, dtype=np.int64)
        B = ids.shape[0]
        optr = np.zeros(B + 1, np.int64); np.cumsum(self.len[ids], out=optr[1:])
        meta = self._ring.stage([ids, optr, np.zeros(B + 1, np.int64)], self.device)
        out = torch.empty(int(optr[-1]), dtype=torch.int32, device=self.device)
        call("tsg_pack_batch_compact", ptr(meta[:B]), ptr(meta[B:2 * B + 1]), ptr(meta[2 * B + 1:]), B, ptr(self.ptr),
             ptr(self._zero_ptr), ptr(self._dummy), ptr(self._dummy), ptr(self.values.view(torch.int32)), ptr(out),
             ptr(self._dummy), ptr(self._dummy), stream_ptr())
        return (out.view(self.values.dtype), optr)


def compact_host_batch(corpus: Corpus, graph_ids, triplets: np.ndarray, pin: bool | None = None,
                       coalesced: bool | None = None) -> dict:
    """One step's HOST batch in the compact form `TripletTrainer.run_from_host_compact` consumes: the chosen graphs'
    node labels (int32), graph-local edge endpoints (int32), node / edge offsets (int64 numpy) and the triplet index
    rows [T, 3] into `graph_ids` (int64).  What a TU file stores, nothing expanded: 4 n + 8 E bytes per graph.
    `pin` (default: when CUDA is available) page-locks the tensors so the feeder's H2D copies are asynchronous."""
    from .synth import select
    sel = select(corpus, np.asarray(graph_ids, dtype=np.int64))
    n_max = int(np.diff(sel.node_ptr).max()) if sel.num_graphs else 0
    if n_max >= 2 ** 31:
        raise ValueError("tsg: graph too large for int32 local node ids")
    trip = np.ascontiguousarray(np.asarray(triplets, dtype=np.int64).reshape(-1, 3))
    if trip.size and (trip.min() < 0 or trip.max() >= sel.num_graphs):
        raise ValueError("tsg: triplet index outside the batch")
    pin = torch.cuda.is_available() if pin is None else pin
    t = lambda a: (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() if pin else torch.from_numpy(np.ascontiguousarray(a)))
    if coalesced is None:
        from .synth import is_coalesced_symmetric
        coalesced = is_coalesced_symmetric(sel)
    return dict(label=t(sel.node_label.astype(np.int32)), row=t(sel.row.astype(np.int32)), col=t(sel.col.astype(np.int32)),
                node_ptr=sel.node_ptr.astype(np.int64).copy(), edge_ptr=sel.edge_ptr.astype(np.int64).copy(),
                triplets=t(trip), coalesced=bool(coalesced))
