#!/bin/bash
# same-box comparison of several builds of libb200ov.so: tools/abn.sh "<pattern>" name1 name2 ...  (tools/ubench/lib_<name>.so)
pat="$1"; shift
cp pyopenvino_b200/libb200ov.so /tmp/lib_orig.so
for v in "$@"; do
  cp tools/ubench/lib_$v.so pyopenvino_b200/libb200ov.so
  echo "== $v"
  timeout 300 python tools/microbench.py --batch ${AB_BATCH:-64} --only "$pat" 2>&1 | grep " ms " | grep -v conv0 | awk '{printf "%s %s %s  %s ms  %s TF\n", $1, $2, $3, $(NF-12), $(NF-1)}'
done
cp /tmp/lib_orig.so pyopenvino_b200/libb200ov.so
