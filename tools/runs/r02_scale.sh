#!/bin/bash
# round 2, 8-GPU call: data-parallel parity test (2 GPUs), 8-GPU bench with the classifier all-reduce overlapped (DG_OVERLAP_AR=1) and not
O=gpurun_out/r02_scale; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -rA -s > $O/pytest_dp.log 2>&1; echo "dp test rc=$?" >> $O/status.txt
N=$(nvidia-smi -L | wc -l)
for ov in 0 1; do
  DG_OVERLAP_AR=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_${N}gpu_overlap$ov.json 2> $O/bench_${N}gpu_overlap$ov.err
  echo "bench N=$N overlap=$ov rc=$?" >> $O/status.txt
done
timeout 300 python bench.py --steps 100 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench N=1 rc=$?" >> $O/status.txt
cat $O/status.txt; tail -4 $O/pytest_dp.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
