#!/bin/bash
# A/B bench on the GPU box: tools/ab.sh NAME1 "ENV1" NAME2 "ENV2" ...   (ENV like "DG_TUNE=9=0")
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  name=$1; envs=$2; shift 2
  env $envs timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$name.json"))
    print("$name", "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]),
          {k: (v["ms"], v["launches"]) for k, v in d["roofline"]["classes"].items() if v["ms"] > 1})
except Exception as e:
    print("$name", "FAILED", e, open("gpurun_out/bench_$name.err").read()[-1500:])
PY
done
