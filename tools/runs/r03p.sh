#!/bin/bash
O=gpurun_out/r03p; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
for rep in 1 2 3; do
  for v in 0 1; do
    DG_TUNE=24=$v timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_split${v}_$rep.json 2> $O/split${v}_$rep.err; echo "bench split=$v $rep rc=$?" >> $O/status.txt
  done
done
cat $O/status.txt | head -3; tail -2 $O/pytest.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
