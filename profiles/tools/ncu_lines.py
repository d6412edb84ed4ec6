import csv, re, sys, subprocess
rep, func, srcfile, idx = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0
lines = open('/tmp/cub/dis.txt').read().split('\n')
start = [i for i, l in enumerate(lines) if ('.text.' + func) in l and 'section' in l][0]
cur = None; seq = []
for l in lines[start:]:
    if l.startswith('//------') and seq: break
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m: cur = int(m.group(2)) if m.group(1).endswith(srcfile) else -1; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
off2line = {o: ln for o, ln, _ in seq}
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
# several kernels: split on "Kernel Name" rows
blocks = []; curb = None
for r in csv.reader(out.split('\n')):
    if r and r[0] == 'Kernel Name': curb = {'name': r[1], 'rows': []}; blocks.append(curb); continue
    if curb is not None: curb['rows'].append(r)
b = blocks[idx]
print(b['name'])
h = b['rows'][0]; ai = h.index('Address'); ns = h.index('# Samples'); ie = h.index('Instructions Executed')
base = None; agg = {}
for r in b['rows'][1:]:
    try: a = int(r[ai], 16)
    except Exception: continue
    if base is None: base = a
    ln = off2line.get(a - base)
    d = agg.setdefault(ln, [0, 0]); d[0] += int(r[ns]); d[1] += int(r[ie])
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
src = open('/root/repo/two-stage-gnn_b200/csrc/' + srcfile).read().split('\n')
print('samples', ts, 'warp-inst', ti)
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 30]:
    s = src[ln - 1].strip()[:105] if ln and ln > 0 else str(ln)
    print(f"{100*v[0]/ts:5.1f}% smp {100*v[1]/ti:5.1f}% inst  L{ln}: {s}")
