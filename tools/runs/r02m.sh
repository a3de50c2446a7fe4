#!/bin/bash
# 2-GPU check of the fused exchange kernel (csrc/dg_dp.cu): parity test, then bench A/B NCCL vs fused (multicast / peer loads)
O=gpurun_out/r02m; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -s > $O/pytest_dp.log 2>&1; echo "dp test rc=$?" >> $O/status.txt
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 200 --warmup 10 --no-cpu-baseline --no-profile > $O/bench_$name.json 2> $O/bench_$name.err
  echo "bench $name rc=$?" >> $O/status.txt
}
run nccl DG_DP_FUSED=0
run fused_mc DG_DP_FUSED=1 DG_DP_MULTICAST=1
run fused_p2p DG_DP_FUSED=1 DG_DP_MULTICAST=0
run nccl_b DG_DP_FUSED=0
run fused_mc_b DG_DP_FUSED=1 DG_DP_MULTICAST=1
cat $O/status.txt; tail -12 $O/pytest_dp.log
grep -h "downgan_b200.dp" $O/*.err | head
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
