#!/bin/bash
O=gpurun_out/r02t; mkdir -p $O
DG_LOG_FALLBACK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_primitives" > $O/pytest_conv.log 2>&1; echo "conv primitives rc=$?" >> $O/status.txt
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg4 --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err; echo "cfg4 rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/pytest_conv.log; grep -h "dg fallback" $O/*.err $O/pytest_conv.log | sort | uniq -c
python - <<PY
import json
d = json.loads(open("$O/cfg4.json").read().strip().splitlines()[-1])
print("cfg4 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
for k, v in sorted(d["roofline"]["classes"].items(), key=lambda kv: -kv[1]["ms"]): print("   ", k, v)
PY
