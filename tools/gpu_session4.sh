#!/bin/bash
# GPU session 4 (round 2): cp.async-staged blur rows + TMA L2 prefetch (default library), default cluster backward.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
V=tools/_build/variants
echo "== timing" > $O/s4_timing.txt
CSPN_B200_LIB=cspn_b200/_build/libcspn_b200.so timeout 300 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 48 1 228 304 24 2>&1 | tail -5 | cut -c1-110 >> $O/s4_timing.txt
echo "== trace" > $O/s4_trace.txt
CSPN_B200_LIB=$V/lib_trace_p4.so timeout 300 python tools/trace_cluster.py >> $O/s4_trace.txt 2>&1
echo "== full GPU suite, default library" > $O/s4_suite.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 >> $O/s4_suite.txt
echo "== train step" > $O/s4_bwd.txt
timeout 200 python tools/time_bwd.py 8 228 304 24 >> $O/s4_bwd.txt 2>&1
timeout 200 python tools/time_bwd.py 4 352 1216 24 >> $O/s4_bwd.txt 2>&1
CSPN_B200_BWD=steps timeout 200 python tools/time_bwd.py 4 352 1216 24 >> $O/s4_bwd.txt 2>&1
echo "== bench" > $O/s4_bench.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline >> $O/s4_bench.txt 2>$O/s4_bench_err.txt
echo done > $O/s4_done.txt
