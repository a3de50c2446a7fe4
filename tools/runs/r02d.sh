#!/bin/bash
# round 2, GPU call D: streaming implicit-GEMM conv kernel - correctness (forced everywhere), A/B per layer, bench A/B
O=gpurun_out/r02d; mkdir -p $O
DG_IG=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_primitives" > $O/pytest_prims_ig2.log 2>&1; echo "prims(ig forced) rc=$?" >> $O/status.txt
DG_IG=2 DG_IG_PLAN=1 timeout 600 python tools/ig_ab.py > $O/ig_ab.md 2> $O/ig_ab.err; echo "ig_ab rc=$?" >> $O/status.txt
DG_IG=2 timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all_ig2.log 2>&1; echo "all(ig forced) rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "all(default) rc=$?" >> $O/status.txt
for v in 0 1; do DG_TUNE=16=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ig$v.json 2> $O/bench_ig$v.err; echo "bench ig$v rc=$?" >> $O/status.txt; done
DG_IG=2 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_ig_forced.json 2> $O/bench_ig_forced.err; echo "bench forced rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/pytest_prims_ig2.log; cat $O/ig_ab.md
