#!/bin/bash
O=gpurun_out/r03b; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
   -k regex:"conv_ws_kernel<\(int\)2, \(int\)1," -s 1 -c 1 -f -o $O/conv_ws_l1dgrad python tools/cycle.py > $O/ncu_full_ws.log 2>&1; echo "ncu ws rc=$?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/ncu_full_ws.log; ls -la $O
