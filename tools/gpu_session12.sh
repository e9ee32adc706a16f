#!/bin/bash
# GPU session 12 (round 2): chained strips as shipped: full GPU suite, timing, bench line, ncu launch list + full capture.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== full GPU suite" > $O/s12_suite.txt
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 2>&1 | tail -8 >> $O/s12_suite.txt
echo "== timing: default" > $O/s12_timing.txt
timeout 120 python tools/time_shape.py cluster 32 352 1216 24 64 228 304 24 64 228 304 48 1 228 304 24 2>&1 | tail -4 | cut -c1-330 >> $O/s12_timing.txt
echo "== bench" > $O/s12_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 >> $O/s12_bench.txt 2>$O/s12_bench_err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launch_list_chained.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/s12_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cspn2d_cluster -s 2 -c 1 -o $O/r02_cluster_chained python tools/run_once.py cluster 3 > $O/s12_ncu.log 2>&1
echo done > $O/s12_done.txt
