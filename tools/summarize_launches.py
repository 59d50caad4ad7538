"""Summarize an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    name = r[ki]
    name = name.split('(')[0] if not name.startswith('void') else name[5:].split('(')[0]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    tot[name[:70]] += v; cnt[name[:70]] += 1
all_ = sum(tot.values())
print(f"total kernel time {all_/1e3:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{v/1e3:10.2f} ms {100*v/all_:5.1f}%  n={cnt[k]:6d}  avg {v/cnt[k]:9.1f} us  {k}")
