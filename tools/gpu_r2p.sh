#!/bin/bash
# GPU session P: (hi, lo)-input staging A/B on SSD (B200OV_F16_STAGE=2), ncu --set full of the new stem / 1x1 group / pool_proj launches
mkdir -p gpurun_out
P=gpurun_out/r2p
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
B200OV_F16_STAGE=2 timeout 600 python -m pytest tests/test_gpu_fusion.py tests/test_gpu_stage_group.py -m gpu -q -x 2>&1 | tail -4
for wl in ssd_mobilenet_v1_coco googlenet-v1; do
  B200OV_F16_STAGE=2 python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}_hlstage.json > ${P}_bench_${wl}_hlstage.json 2> ${P}_bench_${wl}_hlstage.err; echo "bench $wl hlstage rc=$?"
  python bench.py $B --workload $wl --layers-out ${P}_layers_${wl}.json > ${P}_bench_${wl}.json 2> ${P}_bench_${wl}.err; echo "bench $wl rc=$?"
  B200OV_F16_STAGE=2 python bench.py $B --workload $wl > ${P}_bench_${wl}_hlstage2.json 2> ${P}_bench_${wl}_hlstage2.err
  python bench.py $B --workload $wl > ${P}_bench_${wl}_2.json 2> ${P}_bench_${wl}_2.err
done
python - <<'PY'
import json
for wl in ('googlenet-v1', 'ssd_mobilenet_v1_coco'):
    for v in ('', '_hlstage', '_2', '_hlstage2'):
        try:
            d = json.loads(open('gpurun_out/r2p_bench_%s%s.json' % (wl, v)).read().strip().splitlines()[-1])
            print(wl, v or '(default)', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
        except Exception as e:
            print(wl, v, 'FAILED', e)
PY
S="--steps 2 --warmup 3 $B"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16x2 -c 7 -o ${P}_conv7_src -f python bench.py $S > /dev/null 2>&1; echo "ncu conv7 rc=$?"
python tools/ncu_summary.py full ${P}_conv7_src.ncu-rep ${P}_conv7_full.txt
ls -la gpurun_out/
