#!/bin/bash
# GPU session C: FP16 storage parity, full suite, ncu --set full of the TMA pool / dw kernels, bench incl. f16 field
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f16_storage.py -q -x > gpurun_out/r2c_pytest_f16.log 2>&1; echo "pytest(f16) rc=$?"
tail -40 gpurun_out/r2c_pytest_f16.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -15 gpurun_out/r2c_pytest_all.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_max_tma -c 3 -o gpurun_out/r2c_pool_tma_full -f \
  python tools/microbench.py --batch 256 --only maxpool --iters 1 > gpurun_out/r2c_ncu_pool.log 2>&1; echo "ncu pool rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_tma -c 6 -o gpurun_out/r2c_dw_tma_full -f \
  python tools/microbench.py --batch 64 --only dw --iters 1 > gpurun_out/r2c_ncu_dw.log 2>&1; echo "ncu dw rc=$?"
python bench.py --no-secondary --storage f16 > gpurun_out/r2c_bench_f16.json 2> gpurun_out/r2c_bench_f16.err; echo "bench f16 rc=$?"
tail -c 800 gpurun_out/r2c_bench_f16.err
python bench.py --no-secondary --workload ssd_mobilenet_v1_coco --storage f16 > gpurun_out/r2c_bench_ssd_f16.json 2> gpurun_out/r2c_bench_ssd_f16.err; echo "bench ssd f16 rc=$?"
