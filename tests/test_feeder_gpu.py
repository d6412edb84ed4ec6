"""Resident-corpus feeders: DeviceCorpus.from_device / DeviceRagged gathers and the 1 M-graph-style tiled corpus of
bench.py's config 5 (here: a 300-id corpus tiled from 40 base graphs) reproduce what the host packer builds."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from tsg import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_ragged_gather(cuda):
    from tsg.feeder import DeviceRagged
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 9, 50)
    ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    vals = rng.normal(size=int(ptr[-1])).astype(np.float32)
    tab = DeviceRagged(ptr, torch.from_numpy(vals).to(cuda))
    ids = np.array([7, 7, 0, 49, 13, 2], np.int64)
    out, optr = tab.gather(ids)
    ref = np.concatenate([vals[ptr[i]:ptr[i + 1]] for i in ids])
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(optr, np.concatenate([[0], np.cumsum(lens[ids])]))
    for _ in range(6):                                   # the pinned staging ring wraps around
        out2, _ = tab.gather(ids)
        assert torch.equal(out, out2)


def test_config5_resident_corpus_assembles_the_host_packers_batch(cuda):
    """bench.py Config5: shard tiled on the GPU from the base corpus, batch gathered by id -> the same x, CSR, cluster
    labels and final pooling operator the host path (synth.select / eigen_synth.pack_operands) builds."""
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    from tsg import dense, eigen_synth, ops
    c5 = bench.Config5(cuda, num_graphs=300, rank=1, world=2, base_graphs=40)
    assert c5.owned.shape[0] == 150 and c5.corpus.num_graphs == 150
    ids = c5.sample(3, 8)                                # 24 shard indices
    b = c5.assemble(ids)
    base_ids = c5.owned[ids] % 40
    sel = synth.select(c5.base, base_ids)
    pk = synth.pack(sel)
    assert torch.equal(b["x"].cpu(), torch.from_numpy(pk["x"]))
    assert torch.equal(b["el"].edge_index().cpu(), torch.from_numpy(pk["edge_index"]))
    po = eigen_synth.pack_operands(c5.base, c5.opnd, base_ids)
    assert np.array_equal(b["cl"].cpu().numpy(), po["pool"][0][1].astype(np.int32))
    assert np.array_equal(b["cptr"].cpu().numpy(), po["cluster_ptr"])
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(cuda)
    src, dst, w = po["final"][0]
    ref = dense.build_rect_csr(ops.EdgeList(t(src), t(dst), int(src.shape[0])), t(w), len(ids), int(po["cluster_ptr"][-1]))
    got = b["final"][0]
    assert torch.equal(got.rowptr, ref.rowptr) and torch.equal(got.colidx, ref.colidx) and torch.equal(got.val, ref.val)
    c5.world = 1                                         # no process group in this test: the step's collective is skipped
    loss = c5.step(b)
    assert torch.isfinite(loss)
