#!/bin/bash
# same-box A/B of two builds of libb200ov.so (tools/ubench/lib_A.so, lib_B.so): alternate them on the conv micro-benchmarks
cp pyopenvino_b200/libb200ov.so /tmp/lib_orig.so
for rep in 1 2; do
  for v in A B; do
    cp tools/ubench/lib_$v.so pyopenvino_b200/libb200ov.so
    echo "== $v (rep $rep)"
    timeout 300 python tools/microbench.py --batch ${AB_BATCH:-64} --only "${1:-3x3}" 2>&1 | grep conv | grep -v conv0 | awk '{printf "%s %s  %s ms  %s TF\n", $1, $2, $4, $(NF-1)}'
  done
done
cp /tmp/lib_orig.so pyopenvino_b200/libb200ov.so
