#!/bin/bash
# GPU session B of round 2: full parity suite (all failures listed), micro-benchmarks of the new TMA pool / dw kernels and split-K matmul, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_ops.py::test_randomised_parity_sweeps tests/test_gpu_round2.py tests/test_alias.py > gpurun_out/r2b_pytest_new.log 2>&1; echo "pytest(new) rc=$?"
tail -30 gpurun_out/r2b_pytest_new.log
python tools/microbench.py --batch 256 --only maxpool > gpurun_out/r2b_mb_pool.txt 2>&1
python tools/microbench.py --batch 64 --only dw > gpurun_out/r2b_mb_dw.txt 2>&1
python tools/microbench.py --batch 1024 --only matmul > gpurun_out/r2b_mb_mm.txt 2>&1
B200OV_POOL_NO_TMA=1 python tools/microbench.py --batch 256 --only maxpool > gpurun_out/r2b_mb_pool_strip.txt 2>&1
B200OV_DW_NO_TMA=1 python tools/microbench.py --batch 64 --only dw > gpurun_out/r2b_mb_dw_strip.txt 2>&1
cat gpurun_out/r2b_mb_pool.txt gpurun_out/r2b_mb_pool_strip.txt gpurun_out/r2b_mb_dw.txt gpurun_out/r2b_mb_dw_strip.txt gpurun_out/r2b_mb_mm.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -15 gpurun_out/r2b_pytest_all.log
python bench.py --layers-out gpurun_out/r2b_layers_googlenet.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench.err
