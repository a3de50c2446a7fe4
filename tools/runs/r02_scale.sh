#!/bin/bash
# round 2, 8-GPU call: 8-GPU bench with the fused exchange kernel (default, NVSwitch multimem), with plain peer loads and with NCCL
O=gpurun_out/r02_scale; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
N=$(nvidia-smi -L | wc -l)
run() {
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 100 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_${N}gpu_$name.json 2> $O/bench_${N}gpu_$name.err
  echo "bench N=$N $name rc=$?" >> $O/status.txt
}
run fused_multimem DG_DP_FUSED=1 DG_DP_MULTICAST=1
run nccl DG_DP_FUSED=0
run fused_p2p DG_DP_FUSED=1 DG_DP_MULTICAST=0
timeout 300 python bench.py --steps 100 --warmup 10 --no-profile --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench N=1 rc=$?" >> $O/status.txt
cat $O/status.txt; grep -h "downgan_b200.dp" $O/*.err | head -3
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
