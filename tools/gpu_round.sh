#!/bin/bash
# One GPU-box session: smoke, parity tests, bench (both engines), ncu launch list + full captures. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=3 --timeout 90 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
if [ "${SKIP_LAYER:-0}" != "1" ]; then echo "== layer bench"; timeout 600 python tools/layer_bench.py E8P12 1 2>&1 | tail -9 | cut -c1-400; fi
echo "== bench"; timeout 900 python bench.py --steps ${BENCH_STEPS:-256} --warmup 16 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench grouped"; timeout 900 python bench.py --steps 128 --warmup 8 --engine grouped --no-cpu-baseline --no-ref-cuda --no-hf-dropin --no-70b > gpurun_out/bench_grouped.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_grouped.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_reference.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
NCUARGS="--steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-bench --no-ref-cuda --no-hf-dropin --no-70b --prompt-len 8"
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py $NCUARGS > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full decode_step"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_step -s 4 -c 1 -f -o gpurun_out/decode_step_full python bench.py $NCUARGS > gpurun_out/ncu_ds.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ds.log
if [ "${SKIP_UMMA_NCU:-0}" != "1" ]; then
echo "== ncu full umma"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:e8p_umma -s 6 -c 1 -f -o gpurun_out/umma_full python tools/umma_bench.py 128 4096x11008 > gpurun_out/ncu_umma.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_umma.log
fi
fi
if [ "${PREFILL_MODEL:-0}" = "1" ]; then echo "== whole-model prefill"; timeout 300 python tools/prefill_model_bench.py 2>&1 | tail -3 | cut -c1-600; fi
echo done
