#!/bin/bash
# GPU session 16 (round 2): one predicated arrival per warp (expect_tx of 0 bytes elsewhere), chained history through 32-bit shared addresses
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
SH="32 352 1216 24 64 228 304 24 64 228 304 48 64 228 304 4 1 228 304 24"
echo "== full GPU suite" > $O/s16_suite.txt
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 2>&1 | tail -4 >> $O/s16_suite.txt
echo "== timing: default" > $O/s16_timing.txt
timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-110 >> $O/s16_timing.txt
echo "== timing: CSPN_B200_CHAIN=0" >> $O/s16_timing.txt
CSPN_B200_CHAIN=0 timeout 120 python tools/time_shape.py cluster $SH 2>&1 | tail -5 | cut -c1-110 >> $O/s16_timing.txt
