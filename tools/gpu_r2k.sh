#!/bin/bash
# GPU session K: GoogLeNet bench with / without the (hi, lo) contraction edges (same box), full parity suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x 2>&1 | tail -12
python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 --layers-out gpurun_out/r2k_layers_googlenet.json > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
B200OV_NO_HL=1 python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2k_bench_nohl.json 2> gpurun_out/r2k_bench_nohl.err; echo "bench(nohl) rc=$?"
python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2k_bench2.json 2> gpurun_out/r2k_bench2.err
B200OV_NO_HL=1 python bench.py --no-secondary --no-f16 --sustain 0 --cpu-budget 1 > gpurun_out/r2k_bench_nohl2.json 2> gpurun_out/r2k_bench_nohl2.err
python -c "
import json
for f in ('r2k_bench','r2k_bench_nohl','r2k_bench2','r2k_bench_nohl2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['launches_per_step'], d['e2e']['value'])
"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -5 gpurun_out/r2k_pytest_all.log
