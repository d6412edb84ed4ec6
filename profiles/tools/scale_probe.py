"""Where does the N-GPU step lose time?  Per rank: (a) step with no collective, (b) host enqueue time of a step,
(c) step with the all-reduce, (d) the all-reduce alone."""
import os, sys, time
sys.path[:0] = ["/root/repo", "/root/repo/two-stage-gnn_b200"]
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
pin = os.environ.get("PROBE_PIN", "1") == "1"
ncpu = os.cpu_count()
if pin and world > 1:
    per = ncpu // world
    os.sched_setaffinity(0, set(range(lr * per, (lr + 1) * per)))
torch.set_num_threads(2)
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if world > 1:
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
torch.manual_seed(777)
model = tnn.PackedSAGNet(corpus.num_node_labels, 32, 32, 0.5, 0.5).to(dev)
trainer = TripletTrainer(model)
def timed(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(n): fn()
    e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n, (t1 - t0) * 1e3 / n
def local_step():
    model.train(); loss = model.native_step(cb, sel.node_ptr, tidx, 1.5); trainer.opt.step()
a_gpu, a_host = timed(local_step)
b_gpu, b_host = timed(lambda: trainer.step(cb, None, sel.node_ptr, tidx))
flat, _ = model._flat_grads()
def ar():
    if world > 1: dist.all_reduce(flat)
c_gpu, c_host = timed(ar)
msg = f"rank {rank}: nodes {int(sel.node_ptr[-1])} | no-collective step {a_gpu:.3f} ms (host enqueue {a_host:.3f}) | with all-reduce {b_gpu:.3f} (host {b_host:.3f}) | all-reduce alone {c_gpu:.3f} (host {c_host:.3f})"
if world > 1:
    out = [None] * world; dist.all_gather_object(out, msg)
    if rank == 0: print("\n".join(out))
    dist.destroy_process_group()
else:
    print(msg)
