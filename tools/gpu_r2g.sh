#!/bin/bash
# GPU session G: pool -> pool_proj fusion (parity, fused vs separate micro-benchmark, GoogLeNet bench with / without)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x > gpurun_out/r2g_pytest_fusion.log 2>&1; echo "pytest(fusion) rc=$?"
tail -15 gpurun_out/r2g_pytest_fusion.log
python tools/microbench.py --batch 256 --only poolconv > gpurun_out/r2g_mb_poolconv.txt 2>&1
cat gpurun_out/r2g_mb_poolconv.txt
python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 --layers-out gpurun_out/r2g_layers_googlenet.json > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
B200OV_NO_POOL_FUSE=1 python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2g_bench_nofuse.json 2> gpurun_out/r2g_bench_nofuse.err; echo "bench(nofuse) rc=$?"
python -c "
import json
for f in ('r2g_bench','r2g_bench_nofuse'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['launches_per_step'], d['e2e']['value'])
"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -8 gpurun_out/r2g_pytest_all.log
