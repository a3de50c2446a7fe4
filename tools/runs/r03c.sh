#!/bin/bash
# A/B of two builds on ONE box: lib_prev.so / lib_new.so are swapped in as libdowngan_b200.so
O=gpurun_out/r03c; mkdir -p $O
L=downgan_b200/csrc
for rep in 1 2 3; do
  for v in prev new; do
    cp $L/lib_$v.so $L/libdowngan_b200.so
    timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_${v}_$rep.json 2> $O/${v}_$rep.err; echo "bench $v $rep rc=$?" >> $O/status.txt
  done
done
cat $O/status.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
