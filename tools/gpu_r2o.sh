#!/bin/bash
# GPU session O: staged 1x1 (TMA pixel tiles), pixel-group stems, pool-fusion bank-conflict fix: parity, A/B bench lines, micro-bench
mkdir -p gpurun_out
P=gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_stage_group.py tests/test_gpu_fusion.py -m gpu -q -x > ${P}_pytest_new.log 2>&1; echo "pytest(new) rc=$?"; tail -15 ${P}_pytest_new.log
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
for wl in googlenet-v1 ssd_mobilenet_v1_coco; do
  python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}.json 2> ${P}_bench_${wl}.err; echo "bench $wl rc=$?"
  B200OV_F16_STAGE=0 python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}_nostage.json > ${P}_bench_${wl}_nostage.json 2> ${P}_bench_${wl}_nostage.err
  B200OV_F16_STEM_GROUP=1 python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}_nogroup.json > ${P}_bench_${wl}_nogroup.json 2> ${P}_bench_${wl}_nogroup.err
  python bench.py $B --workload $wl > ${P}_bench_${wl}_2.json 2> ${P}_bench_${wl}_2.err
done
python - <<'PY'
import json
for wl in ('googlenet-v1', 'ssd_mobilenet_v1_coco'):
    for v in ('', '_nostage', '_nogroup', '_2'):
        try:
            d = json.loads(open('gpurun_out/r2o_bench_%s%s.json' % (wl, v)).read().strip().splitlines()[-1])
            print(wl, v or '(default)', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
        except Exception as e:
            print(wl, v, 'FAILED', e)
PY
python tools/microbench.py --batch 256 --only 'G ' > ${P}_mb_googlenet.txt 2>&1
python tools/microbench.py --batch 64 --only 'S ' > ${P}_mb_ssd.txt 2>&1
B200OV_F16_STAGE=0 python tools/microbench.py --batch 256 --only 'conv' > ${P}_mb_conv_nostage.txt 2>&1
cat ${P}_mb_googlenet.txt ${P}_mb_ssd.txt
timeout 1500 python -m pytest tests -m gpu -q > ${P}_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -8 ${P}_pytest_all.log
