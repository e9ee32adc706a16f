#!/bin/bash
# GPU session 10 (round 2): chained strips (one-sided halos, history blocks in the workspace) vs the unchained plan.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
SH="32 352 1216 24 64 228 304 24 64 228 304 48 64 228 304 4 64 228 304 8 64 228 304 16 1 228 304 24"
echo "== chained-strip tests" > $O/s10_tests.txt
timeout 900 python -m pytest tests/test_chained_strips_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 >> $O/s10_tests.txt
echo "== timing: default (chains when B*C >= 2 x clusters)" > $O/s10_timing.txt
timeout 300 python tools/time_shape.py cluster $SH 2>&1 | tail -7 | cut -c1-330 >> $O/s10_timing.txt
echo "== timing: CSPN_B200_CHAIN=0" >> $O/s10_timing.txt
CSPN_B200_CHAIN=0 timeout 300 python tools/time_shape.py cluster $SH 2>&1 | tail -7 | cut -c1-130 >> $O/s10_timing.txt
echo "== trace (chained, cfg2)" > $O/s10_trace.txt
[ -f $V/lib_trace_chain.so ] && CSPN_B200_LIB=$V/lib_trace_chain.so timeout 300 python tools/trace_cluster.py >> $O/s10_trace.txt 2>&1
echo "== full GPU suite" > $O/s10_suite.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $O/s10_suite.txt
echo done > $O/s10_done.txt
