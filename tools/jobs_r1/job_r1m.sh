#!/bin/bash
# final sanity of the committed tree: full GPU tests, smoke, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench11.json 2> gpurun_out/bench11.err
tail -1 gpurun_out/bench11.json | cut -c1-330
