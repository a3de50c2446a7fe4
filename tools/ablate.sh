#!/bin/bash
# Live critical-path attribution: bench.py (cfg-2, resident batches) with one kernel family at a time switched off
# (DG_ABLATE, csrc/dg_common.cuh).  The drop in ms/step is what that family costs in the overlapped step.
O=${1:-gpurun_out/ablate}; mkdir -p $O
names=(none wgrad_ws wgrad_l1 trunk_fwd trunk_bwd dense_wgrads conv_l1 classifier conv_ig conv_ws pack_adam colsum)
masks=(0 1 2 4 8 16 32 64 128 256 512 1024)
for i in "${!names[@]}"; do
  DG_ABLATE=${masks[$i]} timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-profile > $O/${names[$i]}.json 2> $O/${names[$i]}.err
  echo "${names[$i]} rc=$?" >> $O/status.txt
done
python - <<PY
import json
base = None
print("| family switched off | ms/step | saved µs/step | share of the live step |")
print("|---|---|---|---|")
for n in "${names[@]}".split():
    try:
        d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print("|", n, "| failed |", e, "|"); continue
    ms = d["ms_per_step"]
    if base is None: base = ms
    print(f"| {n} | {ms:.4f} | {(base - ms) * 1e3:.0f} | {(base - ms) / base * 100:.1f} % |")
PY
