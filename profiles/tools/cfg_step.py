"""one BASELINE config's GPU step alone (for ncu launch lists): python profiles/tools/cfg_step.py <3|4|5> [steps]"""
import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/two-stage-gnn_b200"]
sys.argv_backup = sys.argv[:]
cfg, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
sys.argv = ["bench.py", "--config-steps", str(steps), "--corpus5", "20000"]
import importlib.util, torch
spec = importlib.util.spec_from_file_location("bench", "/root/repo/bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
a = b.parse()
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
kt = lambda fn, reps=1, warm=0: 1.0
out = {"3": b.run_config3, "4": b.run_config4, "5": b.run_config5, "1": b.run_config1}[cfg](a, dev, kt, {}, False)
print(cfg, out["ms_per_step"], out["value"])
