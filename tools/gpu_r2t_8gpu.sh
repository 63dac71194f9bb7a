#!/bin/bash
# GPU session T (8 GPUs of one box): final round-2 build at N = 8 / 4 / 2 / 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29600 bench.py --gpus 8 --no-f16 > gpurun_out/r2t_bench_8gpu.json 2> gpurun_out/r2t_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 300 gpurun_out/r2t_bench_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 --no-f16 --no-secondary --sustain 0 > gpurun_out/r2t_bench_4gpu.json 2> gpurun_out/r2t_bench_4gpu.err; echo "bench4 rc=$?"
timeout 600 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --no-f16 --no-secondary --sustain 0 > gpurun_out/r2t_bench_2gpu.json 2> gpurun_out/r2t_bench_2gpu.err; echo "bench2 rc=$?"
python bench.py --no-f16 --no-secondary --sustain 0 --cpu-budget 1 > gpurun_out/r2t_bench_1gpu.json 2> gpurun_out/r2t_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    d = json.loads(open('gpurun_out/r2t_bench_%dgpu.json' % n).read().strip().splitlines()[-1])
    print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']))
    for k, v in d.get('secondary', {}).items():
        print('   ', k, round(v['value']), 'e2e', round(v['e2e']['value']), 'u8', round(v['e2e_u8']['value']))
PY
