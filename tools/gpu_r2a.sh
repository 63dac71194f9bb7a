#!/bin/bash
# GPU session A of round 2: parity suite, default bench line, H2D ceiling, baseline ncu --set full of the pool / dw strip kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.err
python tools/h2d_ceiling.py > gpurun_out/r2a_h2d_1gpu.json 2>&1; echo "h2d rc=$?"
python tools/microbench.py --batch 256 --only maxpool > gpurun_out/r2a_mb_pool.txt 2>&1
python tools/microbench.py --batch 64 --only dw > gpurun_out/r2a_mb_dw.txt 2>&1
python tools/microbench.py --batch 1024 --only matmul > gpurun_out/r2a_mb_mm.txt 2>&1
cat gpurun_out/r2a_mb_pool.txt gpurun_out/r2a_mb_dw.txt gpurun_out/r2a_mb_mm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_max_strip -c 6 -o gpurun_out/r2a_pool_full -f \
  python tools/microbench.py --batch 256 --only maxpool --iters 1 > gpurun_out/r2a_ncu_pool.log 2>&1; echo "ncu pool rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_strip -c 8 -o gpurun_out/r2a_dw_full -f \
  python tools/microbench.py --batch 64 --only dw --iters 1 > gpurun_out/r2a_ncu_dw.log 2>&1; echo "ncu dw rc=$?"
ls -la gpurun_out | head -40
