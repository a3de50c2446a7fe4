#!/bin/bash
O=gpurun_out/r02s; mkdir -p $O
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err; echo "cfg4 rc=$?" >> $O/status.txt
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg3 --steps 20 --warmup 5 --no-cpu-baseline > $O/cfg3.json 2> $O/cfg3.err; echo "cfg3 rc=$?" >> $O/status.txt
DG_LOG_FALLBACK=1 timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg5.json 2> $O/cfg5.err; echo "cfg5 rc=$?" >> $O/status.txt
cat $O/status.txt; grep -h "dg fallback" $O/*.err | sort | uniq -c
