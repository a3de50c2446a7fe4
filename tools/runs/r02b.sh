#!/bin/bash
# round 2, GPU call B: full GPU suite (new adjacent-row tests, L1-sign pinning, bf16 curves), parity report, bench A/B
O=gpurun_out/r02b; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python tools/parity_report.py > $O/parity_r02.md 2> $O/parity.err; echo "parity rc=$?" >> $O/status.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; grep -E "passed|failed" $O/pytest.log | tail -3
