#!/bin/bash
# GPU session I: ncu --set full of the pool-fused contraction (3a pool+proj shape) to see what bounds it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fusion.py -m gpu -q -x > gpurun_out/r2i_pytest_fusion.log 2>&1; echo "pytest(fusion) rc=$?"
tail -5 gpurun_out/r2i_pytest_fusion.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_f16x2 -c 2 -o gpurun_out/r2i_poolconv_full -f \
  python tools/microbench.py --batch 256 --only "3a/pool+proj fused" --iters 1 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2i_ncu.log
