#!/bin/bash
# GPU session Z: how often does a determinism test of tests/test_gpu_models.py fail, per arrive protocol (elected lane / every thread)?
mkdir -p gpurun_out
P=gpurun_out/r2z2
cp pyopenvino_b200/libb200ov.so /tmp/lib_orig.so
for rep in 1 2 3 4; do
  for v in elected thread; do
    cp tools/ubench/lib_$v.so pyopenvino_b200/libb200ov.so
    timeout 300 python -m pytest tests/test_gpu_f16_storage.py tests/test_gpu_fusion.py tests/test_gpu_models.py -m gpu -q > ${P}_${v}_${rep}.log 2>&1
    echo "$v $rep: $(tail -1 ${P}_${v}_${rep}.log) $(grep -h 'differ in rows\|^FAILED' ${P}_${v}_${rep}.log | head -3 | tr '\n' ' ')"
  done
done
cp /tmp/lib_orig.so pyopenvino_b200/libb200ov.so
