import sys, os, time
sys.path[:0] = ["/root/repo", "/root/repo/two-stage-gnn_b200"]
import numpy as np, torch
from tsg import synth, nn as tnn, ops, _lib
dev = torch.device("cuda:0")
c = synth.tile_corpus(synth.make_corpus("DD", 1168, seed=777), 3504)
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
cb = ops.CompactBatch(t(c.node_label, np.int32), t(c.row, np.int32), t(c.col, np.int32), t(c.node_ptr, np.int64),
                      t(c.edge_ptr, np.int64), c.num_node_labels, int(np.diff(c.edge_ptr).max()))
torch.manual_seed(5)
model = tnn.PackedSAGNet(c.num_node_labels, 32, 32, 0.5, 0.0).to(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
with torch.no_grad():
    for fused in (1, 0):
        _lib.lib.tsg_sag_set_fused(fused)
        for _ in range(3): model(cb, None, c.node_ptr)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps): model(cb, None, c.node_ptr)
        e.record(); torch.cuda.synchronize()
        print("fused" if fused else "unfused", "fwd (encoder + head) ms:", s.elapsed_time(e) / reps)
