#!/bin/bash
# GPU session W: direct C_in = 1 stem kernel (mnist / mnist_bn), 2-D grid row Concat: parity + bench lines
mkdir -p gpurun_out
P=gpurun_out/r2w
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -4 ${P}_pytest_all.log
for wl in mnist_bn mnist ssd_mobilenet_v1_coco; do
  python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}.json 2> ${P}_bench_${wl}.err
  B200OV_NO_C1_DIRECT=1 python bench.py $B --workload $wl > ${P}_bench_${wl}_noc1.json 2> ${P}_bench_${wl}_noc1.err
  python - <<PY
import json
for v in ('', '_noc1'):
    d = json.loads(open('${P}_bench_${wl}' + v + '.json').read().strip().splitlines()[-1])
    print('$wl', v or '(default)', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['model_roofline']['frac'])
PY
done
python - <<'PY'
import json
for wl in ('mnist_bn', 'mnist', 'ssd_mobilenet_v1_coco'):
    a = json.load(open('gpurun_out/r2w_layers_%s.json' % wl))['layers']
    for l in a[:4] + [l for l in a if l['kind'] in ('concat',)]:
        print(wl, l['name'][-40:], l['kind'], round(l['ms'], 4), round(l['roofline_ms'], 4))
PY
