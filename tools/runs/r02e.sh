#!/bin/bash
# round 2, GPU call E: early ring fill (key 17), refined IG policy, fused GP reduce / vectorised Adam; cfg4 / cfg5 with IG + wgrad split
O=gpurun_out/r02e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "all rc=$?" >> $O/status.txt
for v in 0 1; do DG_TUNE=17=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_early$v.json 2> $O/bench_early$v.err; echo "bench early$v rc=$?" >> $O/status.txt; done
DG_TUNE=17=0 timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_early0b.json 2> $O/bench_early0b.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_early1b.json 2> $O/bench_early1b.err
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?" >> $O/status.txt
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?" >> $O/status.txt
DG_LOG_FALLBACK=1 timeout 300 python bench.py --config cfg3 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/pytest_all.log; grep -h "dg fallback" $O/*.err | sort | uniq -c
