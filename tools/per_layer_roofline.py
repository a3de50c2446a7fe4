"""Per-launch roofline of ONE critic iteration (cfg-2: B = 64, 3B = 192 rows) from an ncu launch list taken with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (tools/evidence.sh): every launch of the
iteration is labelled with its layer and pass, its algorithmic FLOPs and bytes (SURVEY.md §8d conventions: each operand once)
and compared with the time the measured HBM / bf16 peaks (MEASURED_PEAKS.json) would allow.
Usage: per_layer_roofline.py launches.csv out.md      (expects the one-launch-per-layer weight-gradient schedule)"""
import csv, json, os, re, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM, TF = PK.get("hbm_gbs", 6544.7), PK.get("bf16_tflops_sustained", 1391.2)
B = 64
# critic layers (critic.py:21-86): Ci, Co, Hin, Hout
LAY = [(2, 16, 128, 128), (16, 16, 128, 64), (16, 32, 64, 64), (32, 32, 64, 32), (32, 64, 32, 32), (64, 64, 32, 16),
       (64, 128, 16, 16), (128, 128, 16, 8)]
FC_IN, FC_H = 8192, 100


def conv(l, nb, kind):
    ci, co, hi, ho = LAY[l]
    fl = 2.0 * nb * ho * ho * co * 9 * ci
    x = nb * hi * hi * ci * (4 if l == 0 else 2)   # the 2-channel fields are fp32
    y = nb * ho * ho * co * 2
    if kind == "fwd":
        by = x + y
    elif kind == "jvp":      # reads v_{l-1}, reads the mask a_{l+1} and overwrites it
        by = x + 2 * y
    elif kind == "dgrad":    # reads dz_{l+1}, the mask a_l, writes dz_l
        by = y + 2 * x
    else:                    # wgrad: reads x and dy
        by = x + y
    return fl, by


def schedule():
    """Launch order of one critic iteration (round 2: the real + fake rows of the layer 0-2 weight gradients are launched
    during the input-gradient chain, the interpolates' rows during the JVP chain)."""
    s = [("build [real;fake;interp]", 0.0, 3 * B * 128 * 128 * 2 * 4 + 2 * B * 128 * 128 * 2 * 4)]
    s += [(f"L{l} fwd (3B)",) + conv(l, 3 * B, "fwd") for l in range(8)]
    s += [("fc1 fwd (3B)", 2.0 * 3 * B * FC_IN * FC_H, 3 * B * FC_IN * 2 + FC_IN * FC_H * 4), ("classifier head", 0.0, 3 * B * FC_H * 12)]
    s += [("fc1 dgrad (3B)", 2.0 * 3 * B * FC_IN * FC_H, 2 * 3 * B * FC_IN * 2 + FC_IN * FC_H * 4)]
    s += [(f"L{l} dgrad (3B)",) + conv(l, 3 * B, "dgrad") for l in range(7, 2, -1)]
    s += [("L2 wgrad (real+fake, 2B)",) + conv(2, 2 * B, "wgrad"), ("L2 dgrad (3B)",) + conv(2, 3 * B, "dgrad")]
    s += [("L1 wgrad (real+fake, 2B)",) + conv(1, 2 * B, "wgrad"), ("L1 dgrad (3B)",) + conv(1, 3 * B, "dgrad")]
    s += [("L0 wgrad (real+fake, 2B)",) + conv(0, 2 * B, "wgrad")]
    ci, co, hi, ho = LAY[0]
    s += [("L0 dgrad (interp, B)", 2.0 * B * hi * hi * 2 * 9 * 16, B * hi * hi * (16 * 2 + 2 * 4))]
    s += [("GP norms + finish (one launch)", 0.0, B * 128 * 128 * 2 * 4), ("GP scale (u)", 0.0, 2 * B * 128 * 128 * 2 * 4)]
    for l in range(8):
        nb = B if l <= 2 else 3 * B
        s += [(f"L{l} wgrad ({'interp, B' if l <= 2 else '3B'})",) + conv(l, nb, "wgrad"), (f"L{l} JVP (B)",) + conv(l, B, "jvp")]
    s += [("fc1 wgrad (3B)", 2.0 * 3 * B * FC_IN * FC_H, 3 * B * FC_IN * 2 + 2 * FC_IN * FC_H * 4),
          ("fc1 JVP (B)", 2.0 * B * FC_IN * FC_H, B * FC_IN * 2 + FC_IN * FC_H * 4), ("fc1 JVP finish", 0.0, 0.0),
          ("small classifier grads", 0.0, 0.0), ("unpack gradients", 0.0, 2 * 1112313 * 4), ("Adam", 0.0, 1112313 * 28),
          ("pack weights (fwd + dgrad + K-major images)", 0.0, 1112313 * 4 * 3 + 292896 * 2 * 2)]
    return s


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, st = r, i + 1
            break
    I = {h: i for i, h in enumerate(hdr)}
    by = OrderedDict()
    for r in rows[st:]:
        if len(r) < len(hdr):
            continue
        n = re.sub(r"\(.*", "", r[I["Kernel Name"]]).replace("dg::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
        e = by.setdefault(r[I["ID"]], {"n": n, "us": 0.0, "dram": 0.0})
        v = float(r[I["Metric Value"]].replace(",", ""))
        sc = {"ns": 1e-3, "us": 1, "ms": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[I["Metric Unit"]], 1)
        if r[I["Metric Name"]] == "gpu__time_duration.sum":
            e["us"] = v * sc
        else:
            e["dram"] += v * sc
    L = list(by.values())
    bi = [i for i, l in enumerate(L) if l["n"].startswith("build_critic_input")]
    step = L[bi[-3]:bi[-2]]  # a critic iteration that is not followed by the generator iteration
    sch = schedule()
    if len(step) != len(sch):
        sys.exit(f"launch list has {len(step)} launches per critic iteration, the schedule table {len(sch)}: update schedule()")
    out = ["One critic iteration (cfg-2, B = 64: 3B = 192 rows through the critic), launch by launch, under ncu (serialised, cold",
           f"cache).  Ideal = max(algorithmic bytes / {HBM:.0f} GB/s, FLOPs / {TF:.0f} TFLOP/s); algorithmic bytes count every operand once.",
           "", "| # | kernel | what | µs | GFLOP | alg MB | DRAM MB | TFLOP/s | alg GB/s | ideal µs | × ideal |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    tot = tot_ideal = 0.0
    for k, (l, (what, fl, byts)) in enumerate(zip(step, sch)):
        ideal = max(byts / (HBM * 1e3), fl / (TF * 1e6))
        tot += l["us"]; tot_ideal += ideal
        out.append(f"| {k} | `{l['n'][:26]}` | {what} | {l['us']:.1f} | {fl/1e9:.2f} | {byts/1e6:.1f} | {l['dram']/1e6:.1f} | "
                   f"{fl/l['us']/1e6:.0f} | {byts/l['us']/1e3:.0f} | {ideal:.1f} | {(l['us']/ideal if ideal > 0.05 else float('nan')):.1f} |")
    out.append(f"| | | **total** | **{tot:.0f}** | | | | | | **{tot_ideal:.0f}** | **{tot/tot_ideal:.1f}** |")
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
