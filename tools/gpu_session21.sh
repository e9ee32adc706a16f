#!/bin/bash
# GPU session 21 (round 2): final state: full GPU suite, smoke, bench (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== full GPU suite" > $O/s21_suite.txt
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -4 >> $O/s21_suite.txt
echo "== smoke" >> $O/s21_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 >> $O/s21_suite.txt
timeout 700 python bench.py --steps 20 --warmup 5 > $O/s21_bench.txt 2> $O/s21_bench_err.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s21_bench_ref.txt 2>> $O/s21_bench_err.txt
