#!/bin/bash
# split-K sweep of the tcgen05 decode+GEMM kernel (GPU box): tools/umma_ksplit_sweep.sh "<M list>" "<NxK list>" [rt]
for ks in 1 2 4 8; do echo "ksplit=$ks rt=${3:-0}"; QUIPB200_OPTIONS=umma_ksplit=$ks,umma_rt=${3:-0} timeout 200 python tools/umma_bench.py ${1:-32,256} ${2:-4096x4096} > /dev/null 2>&1; python -c "
import json
d=json.load(open('gpurun_out/umma_bench.json'))
for r in d: print({k:r[k] for k in ('N','K','M','umma_us','dense_us')})
"; done
