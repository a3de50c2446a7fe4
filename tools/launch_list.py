"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of bench.py: the LAST full 5-step cycle
(5 critic + 1 generator iteration), grouped by kernel.  Usage: launch_list.py launches.csv [out.md]"""
import csv, sys, re
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
I = {h: i for i, h in enumerate(hdr)}
L = []
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("dg::", "").replace("<unnamed>::", "").replace("void ", "")
    L.append((name, float(r[I["Metric Value"]]) / 1000.0, r[I["Grid Size"]]))
# cycles are delimited by the generator step's trunk_bwd_kernel; take the launches between the last two of them
idx = [i for i, l in enumerate(L) if l[0].startswith("trunk_bwd_kernel")]
if len(idx) >= 2 and '--all' not in sys.argv:
    # a cycle = from just after the gen step's final adam (after trunk_bwd) ... simpler: window between consecutive trunk_bwd launches
    L = L[idx[-2]:idx[-1]]
tot = sum(l[1] for l in L)
agg = OrderedDict()
for n, t, g in L:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += t
out = [f"One full 5-step cycle (5 critic + 1 generator iteration, B=64) = {len(L)} launches, {tot/1000:.3f} ms of kernel time under ncu",
       "(cold-cache, serialised: compare shares, not absolutes).", "", "| kernel | launches / 5 steps | total us | avg us | share |", "|---|---|---|---|---|"]
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{n}` | {c} | {t:.1f} | {t/c:.2f} | {100*t/tot:.1f}% |")
text = "\n".join(out)
print(text)
outs = [a for a in sys.argv[2:] if not a.startswith('--')]
if outs:
    open(outs[0], "w").write(text + "\n")
