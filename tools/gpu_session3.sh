#!/bin/bash
# GPU session 3 (round 2): rows requested one task ahead + early publish (default library), full GPU suite, trace, ncu.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== timing" > $O/s3_timing.txt
for v in default; do
  lib=cspn_b200/_build/libcspn_b200.so
  echo "-- $v" >> $O/s3_timing.txt
  CSPN_B200_LIB=$lib timeout 300 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 4 64 228 304 8 64 228 304 16 64 228 304 48 1 228 304 24 2>&1 | tail -8 | cut -c1-110 >> $O/s3_timing.txt
done
echo "== trace" > $O/s3_trace.txt
CSPN_B200_LIB=$V/lib_trace_p3.so timeout 300 python tools/trace_cluster.py >> $O/s3_trace.txt 2>&1
echo "== full GPU suite, default library" > $O/s3_suite.txt
CSPN_B200_TEST_STAGED=1 timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 >> $O/s3_suite.txt
echo "== bench" > $O/s3_bench.txt
timeout 900 python bench.py --steps 20 --warmup 5 >> $O/s3_bench.txt 2>$O/s3_bench_err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cspn2d_cluster -s 2 -c 1 -o $O/r02_cluster_p3 python tools/run_once.py cluster 3 > $O/s3_ncu.log 2>&1
echo done > $O/s3_done.txt
