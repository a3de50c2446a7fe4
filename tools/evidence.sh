#!/bin/bash
# Round evidence on the GPU box: tools/evidence.sh TAG [quick]
#   pytest -m gpu, both bench arms, ncu launch list (durations + DRAM bytes of every launch of one steady-state
#   cycle), and one `--set full` capture of the dominant conv kernel.  Everything lands in gpurun_out/.
TAG=${1:-r01}
MODE=${2:-full}
mkdir -p gpurun_out
if [ "$MODE" = full ]; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_$TAG.txt
  timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
fi
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python tools/cycle.py > gpurun_out/cycle_$TAG.log 2>&1
if [ "$MODE" = full ]; then
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:conv_ws_kernel -s 30 -c 1 -f -o gpurun_out/conv_ws_c30_$TAG python tools/cycle.py > gpurun_out/ncu_full_$TAG.log 2>&1
fi
tail -3 gpurun_out/pytest_$TAG.txt 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
      {k: (v["ms"], v["launches"]) for k, v in d["roofline"]["classes"].items() if v["ms"] > 1})
PY
