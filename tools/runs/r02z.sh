#!/bin/bash
O=gpurun_out/r02z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_masks.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
DG_TUNE=20=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_l1 --csv --log-file $O/l1_1.csv python tools/one_conv.py 192 2 16 128 1 > $O/one_1.log 2>&1
run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_$name.json 2> $O/$name.err; echo "bench $name rc=$?" >> $O/status.txt; }
run planar DG_TUNE=20=1
run im2col DG_TUNE=20=0
run planar_b DG_TUNE=20=1
run im2col_b DG_TUNE=20=0
cat $O/status.txt; tail -2 $O/pytest.log
grep -h "gpu__time_duration" $O/l1_1.csv | awk -F'","' '{print $5, $(NF)}'
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
