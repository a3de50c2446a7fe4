"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`)
of tools/cycle.py or bench.py: one full 5-step cycle (5 critic + 1 generator iteration), grouped by kernel and by
bench.py profile class.  Usage: launch_list.py launches.csv [out.md] [--traffic out.json] [--all]"""
import csv, json, re, sys
from collections import OrderedDict

CLASS_OF = [  # kernel-name prefix -> bench.py roofline class (csrc: prof_begin(PC_*))
    ("conv_ws_kernel", "conv_tcgen05"), ("conv_ig_kernel", "conv_tcgen05"), ("conv_umma_kernel", "conv_tcgen05"), ("conv_l1_kernel", "conv_tcgen05"),
    ("wgrad_ws", "wgrad_tcgen05"), ("wgrad_l1_kernel", "wgrad_tcgen05"), ("wgrad_umma", "wgrad_tcgen05"),
    ("trunk_", "dense_block_tcgen05"), ("fc_", "linear"), ("fc2_", "linear"), ("critic_small_grads", "linear"),
    ("conv_co2", "conv_direct"), ("conv_ci2", "conv_direct"), ("conv_direct", "conv_direct"),
    ("wgrad_skinny", "wgrad_direct"), ("wgrad_direct", "wgrad_direct"),
    ("adam_kernel", "adam"), ("l1_kernel", "l1_loss"), ("gp_", "gp_norm"), ("build_critic_input", "interpolate"),
]


def unit_scale(u):
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
I = {h: i for i, h in enumerate(hdr)}
by_id = OrderedDict()
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    name = re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("dg::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    e = by_id.setdefault(r[I["ID"]], {"name": name, "grid": r[I["Grid Size"]], "us": 0.0, "rd": None, "wr": None})
    v = float(r[I["Metric Value"]].replace(",", "")) * unit_scale(r[I["Metric Unit"]])
    m = r[I["Metric Name"]]
    if m == "gpu__time_duration.sum":
        e["us"] = v
    elif m == "dram__bytes_read.sum":
        e["rd"] = v
    elif m == "dram__bytes_write.sum":
        e["wr"] = v
L = list(by_id.values())
# --cycle: the list holds several cycles (bench.py under ncu), delimited by the generator step's trunk_bwd_kernel launches
# (two per step since the backward runs in RRDB ranges); default / --all: the list IS one cycle (tools/cycle.py)
idx = [i for i, l in enumerate(L) if l["name"].startswith("trunk_bwd_kernel")]
if len(idx) >= 4 and "--cycle" in sys.argv:
    L = L[idx[-4]:idx[-2]]
tot = sum(l["us"] for l in L)
have_dram = any(l["rd"] is not None for l in L)
agg = OrderedDict()
for l in L:
    a = agg.setdefault(l["name"], [0, 0.0, 0.0])
    a[0] += 1; a[1] += l["us"]; a[2] += (l["rd"] or 0) + (l["wr"] or 0)
out = [f"One full 5-step cycle (5 critic + 1 generator iteration, B=64) = {len(L)} launches, {tot/1000:.3f} ms of kernel time under ncu",
       "(cold-cache, serialised: compare shares, not absolutes).", "",
       "| kernel | launches / 5 steps | total us | avg us | share |" + (" DRAM MB / launch | DRAM GB/s |" if have_dram else ""),
       "|---|---|---|---|---|" + ("---|---|" if have_dram else "")]
for n, (c, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{n}` | {c} | {t:.1f} | {t/c:.2f} | {100*t/tot:.1f}% |" + (f" {b/c/1e6:.2f} | {b/t/1e3:.0f} |" if have_dram else ""))
cls = OrderedDict()
for l in L:
    k = next((c for p, c in CLASS_OF if l["name"].startswith(p)), "other")
    a = cls.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += l["us"]; a[2] += (l["rd"] or 0) + (l["wr"] or 0)
out += ["", "| bench.py class | launches | total us | share |" + (" DRAM MB / launch | DRAM GB/s |" if have_dram else ""), "|---|---|---|---|" + ("---|---|" if have_dram else "")]
for k, (c, t, b) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {t:.1f} | {100*t/tot:.1f}% |" + (f" {b/c/1e6:.2f} | {b/t/1e3:.0f} |" if have_dram else ""))
text = "\n".join(out)
print(text)
args = [a for a in sys.argv[2:]]
if "--traffic" in args and have_dram:
    p = args[args.index("--traffic") + 1]
    json.dump({k: {"bytes_per_launch": b / c, "launches": c,
                   "source": f"{sys.argv[1]}: sum of dram__bytes_read.sum + dram__bytes_write.sum over the class's launches in one steady-state cycle / launches"}
               for k, (c, t, b) in cls.items()}, open(p, "w"), indent=1)
    args = [a for j, a in enumerate(args) if a != "--traffic" and (j == 0 or args[j - 1] != "--traffic")]
outs = [a for a in args if not a.startswith("--")]
if outs:
    open(outs[0], "w").write(text + "\n")
