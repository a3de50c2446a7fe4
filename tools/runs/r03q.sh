#!/bin/bash
# re-tune two schedule knobs with the final kernels (same box, two passes)
O=gpurun_out/r03q; mkdir -p $O
for rep in 1 2; do
  for t in "13=3" "13=2" "13=4" "13=5" "10=2" "10=6" "10=0"; do
    DG_TUNE=$t timeout 300 python bench.py --no-cpu-baseline --no-profile > $O/bench_${t/=/_}_$rep.json 2> $O/err.log; echo "bench $t $rep rc=$?" >> $O/status.txt
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
PY
