#!/bin/bash
# GPU session E: parity (all), micro-benchmarks of the restructured TMA pool / dw kernels, full bench line with layers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_all.log 2>&1; echo "pytest(all) rc=$?"
tail -25 gpurun_out/r2e_pytest_all.log
python tools/microbench.py --batch 256 --only maxpool > gpurun_out/r2e_mb_pool.txt 2>&1
python tools/microbench.py --batch 64 --only dw > gpurun_out/r2e_mb_dw.txt 2>&1
python tools/microbench.py --batch 256 --only "G conv1" > gpurun_out/r2e_mb_stem.txt 2>&1
python tools/microbench.py --batch 64 --only "S conv0" >> gpurun_out/r2e_mb_stem.txt 2>&1
cat gpurun_out/r2e_mb_pool.txt gpurun_out/r2e_mb_dw.txt gpurun_out/r2e_mb_stem.txt
python bench.py --layers-out gpurun_out/r2e_layers_googlenet.json > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2e_bench.err
python bench.py --no-secondary --no-f16 --workload ssd_mobilenet_v1_coco --layers-out gpurun_out/r2e_layers_ssd.json > gpurun_out/r2e_bench_ssd.json 2> gpurun_out/r2e_bench_ssd.err; echo "bench ssd rc=$?"
