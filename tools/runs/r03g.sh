#!/bin/bash
# 4-GPU sanity run of the default (fused multimem) exchange with the final code, as the driver launches it
O=gpurun_out/r03g; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo "bench N=$N rc=$?" >> $O/status.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/ref_${N}gpu.json 2> $O/ref_${N}gpu.err; echo "reference arm N=$N rc=$?" >> $O/status.txt
cat $O/status.txt; grep -h "downgan_b200.dp" $O/*.err | head -3
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d.get("impl"), d["n_gpus"], round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("scaling"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
PY
