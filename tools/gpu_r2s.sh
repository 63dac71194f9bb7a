#!/bin/bash
# GPU session S: one mbarrier arrival per warp (pool / staged rings, accumulator-drained barriers): parity + bench + micro-bench
mkdir -p gpurun_out
P=gpurun_out/r2s
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
timeout 900 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_stage_group.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -4
for wl in googlenet-v1 ssd_mobilenet_v1_coco; do
  for i in 1 2; do
    python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}_$i.json 2> ${P}_bench_${wl}_$i.err
    python - <<PY
import json
d = json.loads(open('${P}_bench_${wl}_$i.json').read().strip().splitlines()[-1])
print('$wl', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
PY
  done
done
python tools/microbench.py --batch 256 --only 'G ' > ${P}_mb_googlenet.txt 2>&1
python tools/microbench.py --batch 64 --only 'S ' > ${P}_mb_ssd.txt 2>&1
cat ${P}_mb_googlenet.txt ${P}_mb_ssd.txt | grep -v "lrn\|dw\|maxpool"
timeout 1500 python -m pytest tests -m gpu -q -x > ${P}_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -3 ${P}_pytest_all.log
