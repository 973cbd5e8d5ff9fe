#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest gpu"; timeout 300 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== layer bench"; timeout 300 python tools/layer_bench.py E8P12 1 2>&1 | tail -8 | cut -c1-330
echo "== bench"; timeout 600 python bench.py --steps 128 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
