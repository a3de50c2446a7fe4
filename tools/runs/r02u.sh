#!/bin/bash
O=gpurun_out/r02u; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_masks.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
for v in 0 1 0; do DG_TUNE=3=$v timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_order${v}_$RANDOM.json 2> $O/err$v.log; echo "bench order$v rc=$?" >> $O/status.txt; done
DG_TRUNK_TRACE=1 DG_TUNE=3=0 timeout 120 python tools/cycle.py 2>&1 | grep "trunk trace" | sort | uniq -c | sort -rn | head -4 > $O/trace0.txt
cat $O/status.txt; tail -2 $O/pytest.log; cat $O/trace0.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_order*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
