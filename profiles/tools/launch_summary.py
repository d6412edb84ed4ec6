import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value'); 
recs = []
for r in rows[hi + 2:]:
    if len(r) <= mv: continue
    try: recs.append((r[kn], float(r[mv].replace(',', ''))))
    except Exception: pass
# split steps at k_triplet_fwd
idx = [i for i, (k, _) in enumerate(recs) if 'k_triplet_fwd' in k and 'reduce' not in k]
step = recs[idx[-2]:idx[-1]] if len(idx) >= 2 else recs
agg = collections.OrderedDict()
for k, t in step:
    k = re.sub(r'\(.*', '', k)[:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(t for _, t in step)
print(f"launches {len(step)} total {tot/1000:.1f} us")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t/1000:9.1f} us {100*t/tot:5.1f}%  x{n:2d}  {k}")
