#!/bin/bash
# GPU session 11: chained strips, third cut (per-warp copy-out, deferred announcement): tests, timing, trace
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
SH="32 352 1216 24 64 228 304 24 64 228 304 48 64 228 304 16 32 352 1216 12"
echo "== chained-strip tests" > $O/s11_tests.txt
timeout 150 python -m pytest tests/test_chained_strips_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 40 2>&1 | tail -5 >> $O/s11_tests.txt
echo "== timing: default" > $O/s11_timing.txt
timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-130 >> $O/s11_timing.txt
echo "== timing: CSPN_B200_CHAIN=0" >> $O/s11_timing.txt
CSPN_B200_CHAIN=0 timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-130 >> $O/s11_timing.txt
echo "== trace (chained, cfg2)" > $O/s11_trace.txt
CSPN_B200_LIB=$V/lib_trace_chain.so timeout 120 python tools/trace_cluster.py >> $O/s11_trace.txt 2>&1
