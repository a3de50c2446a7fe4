#!/bin/bash
O=gpurun_out/r02x; mkdir -p $O
for v in 0 1; do
  DG_TUNE=20=$v timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_l1 --csv --log-file $O/l1_$v.csv python tools/one_conv.py 192 2 16 128 1 > $O/one_$v.log 2>&1
  echo "variant $v rc=$?" >> $O/status.txt
done
DG_TUNE=20=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_l1p -s 2 -c 1 -f -o $O/conv_l1p python tools/one_conv.py 192 2 16 128 1 > $O/full.log 2>&1; echo "full rc=$?" >> $O/status.txt
cat $O/status.txt
grep -h "gpu__time_duration\|dram__bytes" $O/l1_0.csv | tail -3
grep -h "gpu__time_duration\|dram__bytes" $O/l1_1.csv | tail -3
