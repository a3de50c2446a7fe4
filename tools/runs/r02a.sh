#!/bin/bash
# round 2, GPU call A: full GPU suite, parity report, fc-umma validation, bench lines per config
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python tools/parity_report.py > $O/parity_r02.md 2> $O/parity.err; echo "parity rc=$?" >> $O/status.txt
DG_TEST_FC_UMMA=1 timeout 300 python -m pytest tests/test_gpu_fc_umma.py -m gpu -q -rA > $O/fc_umma.log 2>&1; echo "fc_umma rc=$?" >> $O/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2_$i.json 2> $O/bench_cfg2_$i.err; echo "bench$i rc=$?" >> $O/status.txt; done
if grep -q "fc_umma rc=0" $O/status.txt; then
  DG_TUNE=14=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg2_fc1.json 2> $O/bench_cfg2_fc1.err; echo "bench fc1 rc=$?" >> $O/status.txt
  DG_TUNE=14=1 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_cfg2_fc1_200.json 2> $O/bench_cfg2_fc1_200.err
fi
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $O/bench_cfg2_200.json 2> $O/bench_cfg2_200.err; echo "bench200 rc=$?" >> $O/status.txt
timeout 400 python bench.py --config cfg3 --steps 40 --warmup 5 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?" >> $O/status.txt
timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?" >> $O/status.txt
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?" >> $O/status.txt
cat $O/status.txt
