#!/bin/bash
# the driver's round-end sequence on a fresh box: GPU tests, smoke, reference arm, default bench
O=gpurun_out/r03k; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $O/ref.json 2> $O/ref.err; echo "reference rc=$?" >> $O/status.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; tail -2 $O/pytest.log; tail -1 $O/smoke.log
python - <<PY
import json
d = json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "gpu_launches", "clocks")})
print(d["e2e"]); print({k: v for k, v in d["roofline"].items() if k != "classes"}); print(d["cpu_baseline"])
PY
