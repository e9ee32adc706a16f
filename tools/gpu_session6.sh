#!/bin/bash
# GPU session 6 (round 2): bank-pinned weights vs not; 3D with L2-resident weights; new tests (metrics, per-channel 3D).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== timing" > $O/s6_timing.txt
for v in default nopin; do
  lib=$V/lib_$v.so; [ $v = default ] && lib=cspn_b200/_build/libcspn_b200.so
  echo "-- $v" >> $O/s6_timing.txt
  CSPN_B200_LIB=$lib timeout 300 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 48 2>&1 | tail -3 | cut -c1-110 >> $O/s6_timing.txt
done
echo "== 3D: does an L2-resident weight set make a step faster?" > $O/s6_3d.txt
for shp in "1 8 96 312" "1 16 96 312" "1 32 96 312" "1 64 96 312" "4 16 96 312" "8 64 96 312"; do
  timeout 200 python tools/time_3d.py $shp 12 2>&1 | grep paddle >> $O/s6_3d.txt
done
echo "== GPU suite" > $O/s6_suite.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $O/s6_suite.txt
echo done > $O/s6_done.txt
