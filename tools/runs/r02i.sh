#!/bin/bash
O=gpurun_out/r02i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "all rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all2.log 2>&1; echo "all (2nd) rc=$?" >> $O/status.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_200.json 2> $O/bench_200.err; echo "bench rc=$?" >> $O/status.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_200b.json 2> $O/bench_200b.err; echo "bench b rc=$?" >> $O/status.txt
timeout 600 python bench.py --config cfg5 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/pytest_all.log; tail -2 $O/pytest_all2.log
python - <<PY
import json
for n in ("bench_200", "bench_200b", "bench_cfg5"):
    d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"])
PY
