#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O
for v in 0 1; do
  DG_TUNE=20=$v timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_l1 --csv --log-file $O/l1_$v.csv python tools/one_conv.py 192 2 16 128 1 > $O/one_$v.log 2>&1
  echo "variant $v rc=$?" >> $O/status.txt
done
run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_$name.json 2> $O/$name.err; echo "bench $name rc=$?" >> $O/status.txt; }
run planar DG_TUNE=20=1
run im2col DG_TUNE=20=0
cat $O/status.txt
grep -h "gpu__time_duration" $O/l1_0.csv | tail -1 | cut -d, -f5,15
grep -h "gpu__time_duration" $O/l1_1.csv | tail -1 | cut -d, -f5,15
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
