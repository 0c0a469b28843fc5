#!/bin/bash
# last run of the round on the final tree: the full GPU suite, smoke(), the default bench line
set -u
mkdir -p gpurun_out
{
echo "== pytest -m gpu (final tree) =="
timeout 2400 python -m pytest tests/ -q -m gpu 2>&1 | tail -6
echo "== smoke =="
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (default flags) =="
timeout 1500 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -1 gpurun_out/bench_final2.json | cut -c1-1200
} > gpurun_out/final_suite.log 2>&1
tail -30 gpurun_out/final_suite.log
