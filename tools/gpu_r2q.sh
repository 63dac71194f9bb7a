#!/bin/bash
# GPU session Q: staged 1x1 with the register-gather row mapping (conflict-free epilogue staging) and two items in flight: parity + A/B
mkdir -p gpurun_out
P=gpurun_out/r2q
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_stage_group.py -m gpu -q -x 2>&1 | tail -4
B200OV_F16_STAGE=2 timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_stage_group.py -m gpu -q -x 2>&1 | tail -4
for wl in googlenet-v1 ssd_mobilenet_v1_coco; do
  for v in 1 0 2 1 0 2; do
    B200OV_F16_STAGE=$v python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}_st$v.json > ${P}_bench_${wl}_st$v.json 2> ${P}_bench_${wl}_st$v.err
    python - <<PY
import json
d = json.loads(open('${P}_bench_${wl}_st$v.json').read().strip().splitlines()[-1])
print('$wl', 'STAGE=$v', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
PY
  done
done
python tools/microbench.py --batch 256 --only 'G ' > ${P}_mb_googlenet.txt 2>&1
python tools/microbench.py --batch 64 --only 'S ' > ${P}_mb_ssd.txt 2>&1
cat ${P}_mb_googlenet.txt ${P}_mb_ssd.txt | grep -v "pool\|lrn\|dw"
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -3 ${P}_pytest_all.log
