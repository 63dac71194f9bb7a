#!/bin/bash
# GPU session U: x4 input layout kernel, one-launch row Concat, batch-wide DetectionOutput top-1 pass: parity + bench lines
mkdir -p gpurun_out
P=gpurun_out/r2u
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -5 ${P}_pytest_all.log
for wl in googlenet-v1 ssd_mobilenet_v1_coco; do
  for i in 1 2; do
    python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}_$i.json 2> ${P}_bench_${wl}_$i.err
    python - <<PY
import json
d = json.loads(open('${P}_bench_${wl}_$i.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), d['launches_per_step'])
PY
  done
done
python - <<'PY'
import json
for wl in ('googlenet-v1', 'ssd_mobilenet_v1_coco'):
    a = json.load(open('gpurun_out/r2u_layers_%s.json' % wl))['layers']
    for l in a:
        if l['kind'] in ('input_layout', 'glue', 'concat', 'sigmoid') and l['ms'] > 0.004:
            print(wl, l['name'][-40:], l['kind'], round(l['ms'], 4), round(l['roofline_ms'], 4))
PY
