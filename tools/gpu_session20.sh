#!/bin/bash
# GPU session 20 (round 2): 3D '26sum' / '26sum_abs' straight from the raw guidance (no prep launch, no weight planes)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== 3D tests" > $O/s20_tests.txt
timeout 400 python -m pytest tests/test_cspn3d_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -8 >> $O/s20_tests.txt
echo "== 3D timing (8x64x96x312, N=12): default library" > $O/s20_3d.txt
timeout 200 python tools/time_3d.py 8 64 96 312 12 2>&1 | tail -2 >> $O/s20_3d.txt
echo "-- CSPN3D_GATHER_BLOCKS=2" >> $O/s20_3d.txt
CSPN_B200_LIB=$V/lib_3d_gb2.so timeout 200 python tools/time_3d.py 8 64 96 312 12 2>&1 | tail -2 >> $O/s20_3d.txt
echo "-- CSPN_B200_3D_PADDLE=planes (prep + weight planes)" >> $O/s20_3d.txt
CSPN_B200_3D_PADDLE=planes timeout 200 python tools/time_3d.py 8 64 96 312 12 2>&1 | tail -2 >> $O/s20_3d.txt
