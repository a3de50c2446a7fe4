#!/bin/bash
O=gpurun_out/r02v; mkdir -p $O
run() { local name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_$name.json 2> $O/$name.err; echo "bench $name rc=$?" >> $O/status.txt; }
run base DG_TUNE=19=0
run two_streams DG_TUNE=19=1
run tmem256 DG_TUNE=19=0 DG_WW_TMEM=256
run two_streams_tmem256 DG_TUNE=19=1 DG_WW_TMEM=256
run base_b DG_TUNE=19=0
DG_TUNE=19=1 timeout 600 python -m pytest tests/test_gpu_masks.py tests/test_gpu_properties.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest (19=1) rc=$?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/pytest.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
