#!/bin/bash
# GPU session 9 (round 2): forward step on FFMA2 with strided column pairs (iterate_p) vs the scalar step (CSPN_NO_PAIRS).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
SH="32 352 1216 24 64 228 304 24 64 228 304 48 64 228 304 4 1 228 304 24"
echo "== 2D tests, default library (pairs)" > $O/s9_tests.txt
timeout 900 python -m pytest tests/test_cspn2d_gpu.py tests/test_cluster_edges_gpu.py tests/test_reference_model_golden_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 >> $O/s9_tests.txt
echo "== timing: pairs (default)" > $O/s9_timing.txt
timeout 300 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-120 >> $O/s9_timing.txt
echo "== timing: scalar step (CSPN_NO_PAIRS)" >> $O/s9_timing.txt
CSPN_B200_LIB=$V/lib_nopairs.so timeout 300 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-120 >> $O/s9_timing.txt
echo "== trace (pairs)" > $O/s9_trace.txt
CSPN_B200_LIB=$V/lib_trace_pairs.so timeout 300 python tools/trace_cluster.py >> $O/s9_trace.txt 2>&1
echo "== full GPU suite" > $O/s9_suite.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $O/s9_suite.txt
echo done > $O/s9_done.txt
