#!/bin/bash
O=gpurun_out/r03h; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -x -s -k "kernel_selection" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; grep "tuning\|passed\|failed\|Error" $O/pytest.log | head -20
