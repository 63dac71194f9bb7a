#!/bin/bash
# GPU session F (8 GPUs of one box): concurrent-H2D ceiling at N = 1/2/4/8 and the 8-GPU bench line (GoogLeNet + secondary workloads)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python tools/h2d_ceiling.py > gpurun_out/r2f_h2d_n1.json 2> gpurun_out/r2f_h2d_n1.err; echo "h2d n1 rc=$?"
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2953$n tools/h2d_ceiling.py > gpurun_out/r2f_h2d_n$n.json 2> gpurun_out/r2f_h2d_n$n.err; echo "h2d n$n rc=$?"
done
cat gpurun_out/r2f_h2d_n*.json
timeout 900 $TR --nproc-per-node 8 --master-port 29600 bench.py --gpus 8 --no-f16 > gpurun_out/r2f_bench_8gpu.json 2> gpurun_out/r2f_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 400 gpurun_out/r2f_bench_8gpu.err
timeout 600 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 --no-f16 --no-secondary --sustain 0 > gpurun_out/r2f_bench_4gpu.json 2> gpurun_out/r2f_bench_4gpu.err; echo "bench4 rc=$?"
nvidia-smi topo -m > gpurun_out/r2f_topo.txt 2>&1; lscpu | head -30 >> gpurun_out/r2f_topo.txt
