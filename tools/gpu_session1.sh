#!/bin/bash
# GPU session 1 (round 2): parity of the new step formulations, A/B timing, per-warp traces, staged backward, 3D ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_smi.txt 2>&1
echo "== parity, default library" > $O/s1_parity.txt
timeout 900 python -m pytest tests/test_cspn2d_gpu.py tests/test_cluster_edges_gpu.py -m gpu -q -x 2>&1 | tail -15 >> $O/s1_parity.txt
for v in step2 step3 step3dep step3dep_sm; do
  echo "== parity, $v" >> $O/s1_parity.txt
  CSPN_B200_LIB=$V/lib_$v.so timeout 900 python -m pytest tests/test_cspn2d_gpu.py tests/test_cluster_edges_gpu.py -m gpu -q -x 2>&1 | tail -6 >> $O/s1_parity.txt
done
echo "== timing" > $O/s1_timing.txt
for v in default step2 step3 step3dep step3dep_sm; do
  lib=$V/lib_$v.so; [ $v = default ] && lib=cspn_b200/_build/libcspn_b200.so
  echo "-- $v" >> $O/s1_timing.txt
  CSPN_B200_LIB=$lib timeout 300 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 4 64 228 304 48 2>&1 | tail -5 >> $O/s1_timing.txt
done
for v in trace1 trace3dep; do
  echo "== trace $v" > $O/s1_$v.txt
  CSPN_B200_LIB=$V/lib_$v.so timeout 300 python tools/trace_cluster.py >> $O/s1_$v.txt 2>&1
done
echo "== staged backward" > $O/s1_bwd.txt
CSPN_B200_TEST_STAGED=1 timeout 600 python -m pytest tests/test_staged_cluster_backward_gpu.py tests/test_backward_gpu.py -m gpu -q 2>&1 | tail -15 >> $O/s1_bwd.txt
timeout 200 python tools/time_bwd.py 8 228 304 24 >> $O/s1_bwd.txt 2>&1
timeout 200 python tools/time_bwd.py 4 352 1216 24 >> $O/s1_bwd.txt 2>&1
CSPN_B200_BWD=cluster timeout 200 python tools/time_bwd.py 8 228 304 24 >> $O/s1_bwd.txt 2>&1
CSPN_B200_BWD=cluster timeout 200 python tools/time_bwd.py 4 352 1216 24 >> $O/s1_bwd.txt 2>&1
echo "== 3D" > $O/s1_3d.txt
timeout 200 python tools/time_3d.py >> $O/s1_3d.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step3d -s 30 -c 1 -o $O/r02_step3d_shipped python tools/time_3d.py 1 64 96 312 12 > $O/s1_ncu3d.log 2>&1
echo done > $O/s1_done.txt
