#!/bin/bash
# GPU session 14 (round 2): task groups of the chained plan (L2 locality of the guidance columns two neighbouring tiles share)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
SH="32 352 1216 24 64 228 304 24 48 352 1216 24"
echo "== chained tests" > $O/s14_tests.txt
timeout 150 python -m pytest tests/test_chained_strips_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 40 2>&1 | tail -3 >> $O/s14_tests.txt
for g in default 0 16 20; do
  echo "== timing: CSPN_B200_CHAIN_GROUP=$g" >> $O/s14_timing.txt
  if [ $g = default ]; then timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -3 | cut -c1-110 >> $O/s14_timing.txt
  else CSPN_B200_CHAIN_GROUP=$g timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -3 | cut -c1-110 >> $O/s14_timing.txt; fi
done
for g in 0 16; do
  CSPN_B200_CHAIN_GROUP=$g CSPN_B200_COOP=0 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cspn2d_cluster -s 2 -c 1 --csv --log-file $O/s14_dram_g$g.csv python tools/run_once.py cluster 3 > /dev/null 2>&1
done
