#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== timeline"; timeout 300 python tools/timeline.py 4096 4096 2>&1 | tail -14
echo "== layer bench"; timeout 600 python tools/layer_bench.py E8P12 1 2>&1 | tail -8 | cut -c1-420
echo "== bench"; timeout 900 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
