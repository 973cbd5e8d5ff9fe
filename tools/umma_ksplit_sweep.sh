for ks in 1 2 4 8; do echo "ksplit=$ks"; QUIPB200_OPTIONS=umma_ksplit=$ks timeout 200 python tools/umma_bench.py 32,256 4096x4096 > /dev/null 2>&1; python -c "
import json
d=json.load(open('gpurun_out/umma_bench.json'))
for r in d: print({k:r[k] for k in ('N','K','M','umma_us','dense_us')})
"; done
