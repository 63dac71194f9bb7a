#!/bin/bash
# GPU session J: GoogLeNet bench with / without the pool -> pool_proj fusion (same box), full parity suite
mkdir -p gpurun_out
python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 --layers-out gpurun_out/r2j_layers_googlenet.json > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
B200OV_NO_POOL_FUSE=1 python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2j_bench_nofuse.json 2> gpurun_out/r2j_bench_nofuse.err; echo "bench(nofuse) rc=$?"
python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2j_bench2.json 2> gpurun_out/r2j_bench2.err
B200OV_NO_POOL_FUSE=1 python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2j_bench_nofuse2.json 2> gpurun_out/r2j_bench_nofuse2.err
python -c "
import json
for f in ('r2j_bench','r2j_bench_nofuse','r2j_bench2','r2j_bench_nofuse2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['launches_per_step'], d['e2e']['value'])
"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -5 gpurun_out/r2j_pytest_all.log
