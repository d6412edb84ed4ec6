"""raw ncu CSV (--page raw --csv of a --set full capture) -> one markdown row per kernel (largest launch of each name)."""
import csv, sys, re
M = {"dur": "gpu__time_duration.sum", "dr": "dram__bytes_read.sum", "dw": "dram__bytes_write.sum",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1hit": "l1tex__t_sector_hit_rate.pct",
     "l2hit": "lts__t_sector_hit_rate.pct", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "inst": "smsp__inst_executed.sum",
     "regs": "launch__registers_per_thread", "smem": "launch__shared_mem_per_block_dynamic",
     "tensor": "sm__inst_executed_pipe_tensor.sum", "tpipe": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "mio_throttle", "lg_throttle", "wait", "math_pipe_throttle", "membar", "not_selected", "branch_resolving", "no_instruction", "dispatch_stall", "drain", "imc_miss", "sleeping", "tex_throttle"]
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
def load(path):
    rows = list(csv.reader(open(path)))
    h = rows[0]; units = rows[1]
    idx = {n: i for i, n in enumerate(h)}
    out = {}
    for r in rows[2:]:
        if len(r) < len(h): continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("tsg::", "")
        d = {k: f(r[idx[m]]) if m in idx else None for k, m in M.items()}
        d["units"] = {k: units[idx[m]] if m in idx else "" for k, m in M.items()}
        st = {s: f(r[idx[STALL % s]]) for s in STALLS if (STALL % s) in idx}
        d["stalls"] = sorted(((v, s) for s, v in st.items() if v), reverse=True)[:3]
        d["grid"] = r[idx["Grid Size"]] if "Grid Size" in idx else ""
        d["block"] = r[idx["Block Size"]] if "Block Size" in idx else ""
        if name not in out or (d["dur"] or 0) > (out[name]["dur"] or 0): out[name] = d
    return out
def us(d):
    v, u = d["dur"], d["units"]["dur"]
    return v / 1000 if u == "ns" else (v if u in ("us", "usecond") else v * 1000 if u == "ms" else v)
def mb(v, u):
    if v is None: return 0
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
if __name__ == "__main__":
    PEAK = 6539.2
    print("| kernel | grid x block | regs | time | DRAM read + write | DRAM GB/s (% of 6,539) | L1 hit | L2 hit | issue active | warps active | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        for name, d in sorted(load(path).items(), key=lambda kv: -(kv[1]["dur"] or 0)):
            t = us(d)
            tr = mb(d["dr"], d["units"]["dr"]) + mb(d["dw"], d["units"]["dw"])
            gbs = tr / t * 1e3 if t else 0
            st = ", ".join(f"{s} {v:.1f}" for v, s in d["stalls"])
            print(f"| `{name[:60]}` | {d['grid']} x {d['block']} | {int(d['regs'] or 0)} | {t:.1f} us | {tr:.0f} MB | {gbs:,.0f} ({100*gbs/PEAK:.0f} %) | "
                  f"{d['l1hit'] or 0:.0f} % | {d['l2hit'] or 0:.0f} % | {d['issue'] or 0:.0f} % | {d['occ'] or 0:.0f} % | {st} |")
