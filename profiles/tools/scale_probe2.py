import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/two-stage-gnn_b200"]
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from tsg import synth, nn as tnn, ops
from tsg.train import TripletTrainer
corpus = synth.make_corpus("DD", 1168, seed=777 + 1_000_003 * rank)
T = 1168
trip = synth.sample_triplets(corpus.y, T, seed=rank)
ids = np.concatenate([trip[:, 0], trip[:, 1], trip[:, 2]])
sel = synth.select(corpus, ids)
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)
cb = ops.CompactBatch(t(sel.node_label, np.int32), t(sel.row, np.int32), t(sel.col, np.int32), t(sel.node_ptr, np.int64),
                      t(sel.edge_ptr, np.int64), corpus.num_node_labels, int(np.diff(sel.edge_ptr).max()), True)
tidx = torch.from_numpy(np.stack([np.arange(T), T + np.arange(T), 2 * T + np.arange(T)], 1).astype(np.int64)).to(dev)
model = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.5).to(dev)
trainer = TripletTrainer(model)
model.train()
acc = {}
def tick(name, t0):
    t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
for it in range(45):
    if it == 5:
        acc.clear(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); T0 = time.perf_counter()
    t0 = time.perf_counter()
    loss = model.native_step(cb, sel.node_ptr, tidx, 1.5); t0 = tick("native_step", t0)
    flat, _ = model._flat_grads(); flat[-1] = 1.0; t0 = tick("fill", t0)
    flat.mul_(float(T)); t0 = tick("mul", t0)
    dist.all_reduce(flat); t0 = tick("all_reduce", t0)
    flat.div_(flat[-1].clone()); t0 = tick("div", t0)
    loss = loss.clone(); t0 = tick("clone", t0)
    trainer.opt.step(); t0 = tick("adam", t0)
torch.cuda.synchronize(); total = (time.perf_counter() - T0) / 40 * 1e3
print(f"rank {rank}: {total:.3f} ms/step wall | host per step: " + " ".join(f"{k} {v/40*1e3:.3f}" for k, v in acc.items()))
dist.destroy_process_group()
