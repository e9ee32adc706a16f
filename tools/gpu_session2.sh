#!/bin/bash
# GPU session 2 (round 2): pipelined prologue, two warp groups with a phase shift; staged backward parity; torch op.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== timing" > $O/s2_timing.txt
for v in default p1s3 p2s3 p2s3g0 p2s3g p2s3g500; do
  lib=$V/lib_$v.so; [ $v = default ] && lib=cspn_b200/_build/libcspn_b200.so
  echo "-- $v" >> $O/s2_timing.txt
  CSPN_B200_LIB=$lib timeout 300 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 4 64 228 304 48 1 228 304 24 2>&1 | tail -6 | cut -c1-110 >> $O/s2_timing.txt
done
for v in p2s3 p2s3g; do
  echo "== parity, $v" >> $O/s2_parity.txt
  CSPN_B200_LIB=$V/lib_$v.so timeout 900 python -m pytest tests/test_cspn2d_gpu.py tests/test_cluster_edges_gpu.py -m gpu -q 2>&1 | tail -12 >> $O/s2_parity.txt
done
for v in trace_p2s3 trace_p2s3g; do
  echo "== trace $v" > $O/s2_$v.txt
  CSPN_B200_LIB=$V/lib_$v.so timeout 300 python tools/trace_cluster.py >> $O/s2_$v.txt 2>&1
done
echo "== full GPU suite, default library" > $O/s2_suite.txt
CSPN_B200_TEST_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 >> $O/s2_suite.txt
echo done > $O/s2_done.txt
