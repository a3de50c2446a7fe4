#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "all rc=$?" >> $O/status.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_200.json 2> $O/bench_200.err; echo "bench rc=$?" >> $O/status.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $O/launches.csv python tools/cycle.py > $O/cycle.log 2>&1; echo "ncu list rc=$?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/pytest_all.log
python - <<PY
import json
for n in ("bench_200",):
    d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["gpu_launches"])
PY
