#!/bin/bash
# GPU session 8 (round 2): 3D 'paddle' step straight from the raw guidance: register budget / hoisting variants, parity, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== 3D timing (tools/time_3d.py 8 64 96 312 12)" > $O/s8_3d.txt
for v in default 3d_pb4 3d_pb2 3d_nohoist; do
  lib=$V/lib_$v.so; [ $v = default ] && lib=cspn_b200/_build/libcspn_b200.so
  echo "-- $v" >> $O/s8_3d.txt
  CSPN_B200_LIB=$lib timeout 200 python tools/time_3d.py 8 64 96 312 12 2>&1 | tail -2 >> $O/s8_3d.txt
done
echo "-- default, CSPN_B200_3D_PADDLE=planes (prep + weight planes, the round-1 path)" >> $O/s8_3d.txt
CSPN_B200_3D_PADDLE=planes timeout 200 python tools/time_3d.py 8 64 96 312 12 2>&1 | tail -2 >> $O/s8_3d.txt
echo "== 3D tests" > $O/s8_tests.txt
timeout 900 python -m pytest tests/test_cspn3d_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 >> $O/s8_tests.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step3d_paddle -s 14 -c 1 -o $O/r02_step3d_paddle python tools/time_3d.py 8 64 96 312 12 > $O/s8_ncu.log 2>&1
echo done > $O/s8_done.txt
