#!/bin/bash
# Round evidence on the GPU box: tools/evidence.sh TAG [quick]
#   pytest -m gpu, both bench arms, ncu launch list (durations + DRAM bytes of every launch of one steady-state
#   cycle), and `--set full` captures of the two tcgen05 conv kernels.  Everything lands in gpurun_out/evidence_TAG/.
TAG=${1:-r02}
MODE=${2:-full}
O=gpurun_out/evidence_$TAG; mkdir -p $O
if [ "$MODE" = full ]; then
  timeout 900 python -m pytest tests -m gpu -q -rA > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
  timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?" >> $O/status.txt
fi
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_20.json 2> $O/bench_20.err; echo "bench20 rc=$?" >> $O/status.txt
timeout 300 python bench.py > $O/bench_200.json 2> $O/bench_200.err; echo "bench200 rc=$?" >> $O/status.txt
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file $O/launches.csv python tools/cycle.py > $O/cycle.log 2>&1; echo "ncu list rc=$?" >> $O/status.txt
if [ "$MODE" = full ]; then
  # full captures: the longest launch of the dominant class (critic layer 1 data gradient: conv_ws_kernel<2, 1>), the streaming
  # conv kernel and the fused trunk forward
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:"conv_ws_kernel<\(int\)2, \(int\)1," -s 1 -c 1 -f -o $O/conv_ws_l1dgrad python tools/cycle.py > $O/ncu_full_ws.log 2>&1; echo "ncu ws rc=$?" >> $O/status.txt
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:conv_ig_kernel -s 2 -c 1 -f -o $O/conv_ig_c2 python tools/cycle.py > $O/ncu_full_ig.log 2>&1; echo "ncu ig rc=$?" >> $O/status.txt
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:trunk_fwd_kernel -s 0 -c 1 -f -o $O/trunk_fwd python tools/cycle.py > $O/ncu_full_trunk.log 2>&1; echo "ncu trunk rc=$?" >> $O/status.txt
  for c in cfg3 cfg4 cfg5; do
    timeout 600 python bench.py --config $c > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?" >> $O/status.txt
  done
fi
cat $O/status.txt
python - <<PY
import json
for n in ("bench_20", "bench_200"):
    d = json.loads(open("$O/%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
          {k: (v["ms"], v["launches"]) for k, v in d["roofline"]["classes"].items() if v["ms"] > 1})
PY
