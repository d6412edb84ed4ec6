"""Helper process of tests/test_launcher.py (CPU box with /root/reference present).  Imports torch / sklearn once, then
forks one child per case and runs them CONCURRENTLY (one thread each); prints one JSON list on stdout.

mode "reach": the real launcher path (tsg.run.install + patch_dense_modules + run_script) with `.cuda()` neutralised
              (there is no GPU in this container): the UNMODIFIED script must get through argument parsing, data loading,
              preprocessing and model construction and into its first forward, which lands in a tsg drop-in and raises
              "no CPU path" -- the product refuses CPU tensors.  Reports which classes of the script's own modules carry
              the B2 patches at that moment.
mode "e2e":   the launcher's compat layer alone (no B2 patches; the sag scripts over the oracle-backed torch_geometric
              of oracle/pyg_oracle_shim): the whole script must run to its last print on CPU with the reference's math.
"""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "two-stage-gnn_b200")
REF = "/root/reference/Code"
sys.path[:0] = [ROOT, PKG]


def cases():
    out = []
    for s in ("train", "train_triplet", "train_triplet_pre_train"):
        out.append(("sag", s, "", ["--dataset=SYN", "--epochs=1", "--nhid=32", "--final_dim=32", "--batch_size=1", "--iterations=1"]))
        for m in ("base", "GAT", "soft-assign"):
            mn = "--max_nodes=100" if s == "train" else "--max-nodes=100"
            out.append(("sage+gat+diffpool", s, m, ["--bmname=SYN", "--datadir=data", "--num_epochs=1", "--input-dim=8",
                                                    "--hidden-dim=8", "--output-dim=8", "--num-classes=2", f"--method={m}", mn]))
        out.append(("eigengcn", s, "", ["--bmname=SYN", "--datadir=data", "--epochs=1", "--hidden-dim=8", "--output-dim=8",
                                        "--num-classes=2", "--max-nodes=100"]))
    return out


def child(mode, case, workdir, data_src):
    import torch
    d, s, m, argv = case
    torch.set_num_threads(1)
    cwd = os.path.join(workdir, f"{mode}-{d.replace('+', '_')}-{s}-{m or 'x'}")
    os.makedirs(os.path.join(cwd, "data"), exist_ok=True)
    os.symlink(data_src, os.path.join(cwd, "data", "SYN"))
    os.chdir(cwd)
    log = open("log.txt", "w")
    os.dup2(log.fileno(), 1); os.dup2(log.fileno(), 2)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    from tsg import run
    script = os.path.join(REF, d, s + ".py")
    res = dict(dir=d, script=s, method=m, mode=mode, ok=False, error="", where="", patched={}, fixes=[])
    try:
        res["install"] = run.install(os.path.dirname(script))
        if mode == "reach":
            done = run.patch_dense_modules(os.path.dirname(script))
            res["patched"] = {k: v[1] for k, v in done.items()}
        elif d == "sag":
            sys.path.insert(0, os.path.join(ROOT, "oracle", "pyg_oracle_shim"))
        g = run.run_script(script, argv)
        res["ok"] = True
        res["fixes"] = g.get("__tsg_fixes__", [])
    except BaseException as e:          # SystemExit from argparse included
        res["error"] = f"{type(e).__name__}: {e}"
        tb = traceback.extract_tb(e.__traceback__)
        res["where"] = [f"{os.path.basename(os.path.dirname(f.filename))}/{os.path.basename(f.filename)}:{f.name}" for f in tb]
    if mode == "reach":
        # are the script's OWN classes (the modules it imported by name) patched right now?
        from tsg import dense_patch
        chk = {}
        enc = sys.modules.get("encoders")
        if enc is not None:
            chk["encoders.GraphConv.forward"] = enc.GraphConv.forward is dense_patch.graphconv_forward
            chk["encoders.GcnEncoderGraph.apply_bn"] = enc.GcnEncoderGraph.apply_bn is dense_patch.apply_bn
            if hasattr(enc, "Pool"):
                chk["encoders.Pool.forward"] = enc.Pool.forward is dense_patch.pool_forward
        gat = sys.modules.get("encoders_GAT")
        if gat is not None:
            chk["encoders_GAT.DGATHead.forward"] = gat.DGATHead.forward is dense_patch.dgathead_forward
        net = sys.modules.get("network")
        if net is not None:
            import tsg.nn
            chk["network.GCNConv is tsg.nn.GCNConv"] = net.GCNConv is tsg.nn.GCNConv
        res["script_classes_patched"] = chk
    log.flush()
    res["tail"] = open("log.txt").read()[-1500:]
    return res


def main():
    mode, workdir = sys.argv[1], sys.argv[2]
    only = set(sys.argv[3:])
    import numpy as np
    import torch  # noqa: F401  (imported once, before the forks; no compute in the parent)
    import sklearn.cluster, sklearn.neighbors, networkx, scipy.sparse  # noqa: F401,E401
    from tsg import synth, tu
    c = synth.make_corpus("PROTEINS", 40, seed=3)
    c.y[:] = np.arange(40) % 2
    tu.write(c, os.path.join(workdir, "src"), "SYN")
    data_src = os.path.join(workdir, "src", "SYN")
    todo = [cs for cs in cases() if not only or cs[0] in only]
    pipes = []
    for cs in todo:
        r, w = os.pipe()
        pid = os.fork()
        if pid == 0:
            os.close(r)
            try:
                res = child(mode, cs, workdir, data_src)
            except BaseException as e:
                res = dict(dir=cs[0], script=cs[1], method=cs[2], mode=mode, ok=False, error=f"driver: {type(e).__name__}: {e}", where=[], tail="")
            os.write(w, json.dumps(res).encode())
            os._exit(0)
        os.close(w)
        pipes.append((pid, r))
    out = []
    for pid, r in pipes:
        buf = b""
        while True:
            chunk = os.read(r, 65536)
            if not chunk:
                break
            buf += chunk
        os.waitpid(pid, 0)
        out.append(json.loads(buf.decode()) if buf else dict(ok=False, error="child died", where=[], tail=""))
    sys.__stdout__.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
