#!/bin/bash
O=gpurun_out/r02j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "all rc=$?" >> $O/status.txt
for v in 1 2 4 2 1; do DG_TUNE=18=$v timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_split${v}_$RANDOM.json 2> $O/err.log; echo "bench split$v rc=$?" >> $O/status.txt; done
cat $O/status.txt; tail -2 $O/pytest_all.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_split*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
