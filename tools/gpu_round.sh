#!/bin/bash
# One GPU-box session: smoke, parity tests, microbench, bench, ncu launch list. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=8 --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
if [ "${SKIP_MICRO:-0}" != "1" ]; then
echo "== microbench"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && timeout 120 /tmp/microbench > gpurun_out/microbench.txt 2>&1; cat gpurun_out/microbench.txt
fi
if [ "${SKIP_LAYER:-0}" != "1" ]; then echo "== layer bench"; timeout 600 python tools/layer_bench.py E8P12 1 2>&1 | tail -9; fi
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-128} --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-kernel-bench --no-ref-cuda --prompt-len 8 > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bench.log
fi
echo done
