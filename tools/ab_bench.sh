#!/bin/bash
# same-box A/B of library builds on whole-model benches: tools/ab_bench.sh name1 name2 ...  (tools/ubench/lib_<name>.so)
cp pyopenvino_b200/libb200ov.so /tmp/lib_orig.so
for rep in 1 2; do
  for v in "$@"; do
    cp tools/ubench/lib_$v.so pyopenvino_b200/libb200ov.so
    for wl in ${AB_WORKLOADS:-googlenet-v1 ssd_mobilenet_v1_coco}; do
      python bench.py --workload $wl --no-secondary --no-f16 --sustain 0 --cpu-budget 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$wl', round(d['value']), round(d['ms_per_step'],4))"
    done
  done
done
cp /tmp/lib_orig.so pyopenvino_b200/libb200ov.so
