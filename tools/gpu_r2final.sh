#!/bin/bash
# GPU session: the committed end-of-round build -- full parity suite twice, smoke(), bench lines of every workload
mkdir -p gpurun_out
P=gpurun_out/r2final
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q > ${P}_pytest_$i.log 2>&1; echo "pytest $i rc=$? $(tail -1 ${P}_pytest_$i.log)"; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
for wl in ssd_mobilenet_v1_coco mnist_bn mnist; do python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}.json 2> ${P}_bench_${wl}.err; done
python bench.py --impl reference > ${P}_bench_ref.json 2> ${P}_bench_ref.err
python - <<'PY'
import json
for f in ('bench', 'bench_ssd_mobilenet_v1_coco', 'bench_mnist_bn', 'bench_mnist', 'bench_ref'):
    d = json.loads(open('gpurun_out/r2final_%s.json' % f).read().strip().splitlines()[-1])
    print(f, round(d['value'], 1), round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1), 'u8', round(d.get('e2e_u8', {}).get('value', 0)))
PY
