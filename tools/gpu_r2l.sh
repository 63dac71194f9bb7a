#!/bin/bash
# GPU session L: programmatic dependent launch (B200OV_PDL=1) -- parity suite and same-box bench A/B
mkdir -p gpurun_out
B200OV_PDL=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest_pdl.log 2>&1; echo "pytest(pdl) rc=$?"
tail -6 gpurun_out/r2l_pytest_pdl.log
for rep in 1 2; do
for v in 0 1; do
  for wl in googlenet-v1 ssd_mobilenet_v1_coco mnist_bn mnist; do
  B200OV_PDL=$v python bench.py --workload $wl --no-secondary --no-f16 --sustain 0 --cpu-budget 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl=$v', '$wl', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"
  done
done
done
