#!/bin/bash
# GPU session AA: (hi, lo) output of the C_in = 1 direct kernel (mnist_bn conv2d -> conv2d_1): parity + bench
mkdir -p gpurun_out
P=gpurun_out/r2aa
timeout 900 python -m pytest tests -m gpu -q -x > ${P}_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 ${P}_pytest.log)"; grep -h "^FAILED\|Error" ${P}_pytest.log | head -5
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
python bench.py $B --workload mnist_bn --layers-out ${P}_layers_mnist_bn.json > ${P}_bench_mnist_bn.json 2> ${P}_bench_mnist_bn.err
B200OV_NO_HL=1 python bench.py $B --workload mnist_bn > ${P}_bench_mnist_bn_nohl.json 2> ${P}_bench_mnist_bn_nohl.err
python - <<'PY'
import json
for v in ('', '_nohl'):
    d = json.loads(open('gpurun_out/r2aa_bench_mnist_bn%s.json' % v).read().strip().splitlines()[-1])
    print('mnist_bn', v or '(default)', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['model_roofline']['frac'])
for l in json.load(open('gpurun_out/r2aa_layers_mnist_bn.json'))['layers'][:6]:
    print(l['name'][-40:], l['kind'], round(l['ms'], 4), round(l['roofline_ms'], 4))
PY
