#!/bin/bash
# GPU session 13 (round 2): ncu --set full of the shipped (chained) cluster kernel (plain launch: CSPN_B200_COOP=0), GPU suite, smoke
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
CSPN_B200_COOP=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:cspn2d_cluster -s 2 -c 1 -o $O/r02_cluster_chained python tools/run_once.py cluster 3 > $O/s13_ncu.log 2>&1
echo "== full GPU suite" > $O/s13_suite.txt
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 60 2>&1 | tail -4 >> $O/s13_suite.txt
echo "== smoke" >> $O/s13_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $O/s13_suite.txt 2>&1
echo done > $O/s13_done.txt
