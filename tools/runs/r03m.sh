#!/bin/bash
# final code on an 8-GPU box: 2-GPU parity test, 8-GPU and 1-GPU bench (default exchange = fused multimem)
O=gpurun_out/r03m; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -s > $O/pytest_dp.log 2>&1; echo "dp test rc=$?" >> $O/status.txt
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench N=$N rc=$?" >> $O/status.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench N=1 rc=$?" >> $O/status.txt
cat $O/status.txt; tail -4 $O/pytest_dp.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
