#!/bin/bash
# GPU session R (final build of round 2): round-2 evidence of the current build: parity suite, default bench line (+ layer tables, reference arm),
# ncu launch lists (GoogLeNet / SSD), ncu --set full of the contraction launches of one GoogLeNet pass, of the TMA pool /
# depthwise tile kernels, and micro-bench sweeps.  ncu runs only after the same command exited 0 without it.
mkdir -p gpurun_out
P=gpurun_out/r2v
timeout 1500 python -m pytest tests -m gpu -q > ${P}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest.log
python bench.py --layers-out ${P}_layers_googlenet.json > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
python bench.py --impl reference > ${P}_bench_ref.json 2> ${P}_bench_ref.err; echo "bench(ref) rc=$?"
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
python bench.py $B --workload ssd_mobilenet_v1_coco --layers-out ${P}_layers_ssd.json > ${P}_bench_ssd.json 2> ${P}_bench_ssd.err; echo "bench(ssd) rc=$?"
python bench.py $B --workload mnist > ${P}_bench_mnist.json 2> ${P}_bench_mnist.err; python bench.py $B --workload mnist_bn --layers-out ${P}_layers_mnist_bn.json > ${P}_bench_mnist_bn.json 2> ${P}_bench_mnist_bn.err; echo "bench(mnist_bn) rc=$?"
python tools/microbench.py --batch 256 --only 'G ' > ${P}_mb_googlenet.txt 2>&1
python tools/microbench.py --batch 64 --only 'S ' > ${P}_mb_ssd.txt 2>&1
S="--steps 2 --warmup 3 $B"
python bench.py $S > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file ${P}_launches_googlenet.csv python bench.py $S > ${P}_ncu_g.log 2>&1; echo "ncu launches(g) rc=$?"
python bench.py $S --workload ssd_mobilenet_v1_coco > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file ${P}_launches_ssd.csv python bench.py $S --workload ssd_mobilenet_v1_coco > ${P}_ncu_s.log 2>&1; echo "ncu launches(ssd) rc=$?"
# full captures: the 40 contraction launches of the first batch-256 pass, the 4 stride-2 pools, the 13 depthwise layers of SSD
timeout 900 ncu --set full --clock-control none -k regex:conv_f16x2 -c 40 -o ${P}_conv_full -f python bench.py $S > ${P}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pool_max_tma -c 4 -o ${P}_pool_full -f python bench.py $S > ${P}_ncu_pool.log 2>&1; echo "ncu pool rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_tma -c 13 -o ${P}_dw_full -f python bench.py $S --workload ssd_mobilenet_v1_coco > ${P}_ncu_dw.log 2>&1; echo "ncu dw rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'lrn_vec4|nchw_to_nhwc|concat_rows|detection' -c 4 -o ${P}_misc_full -f python bench.py $S > ${P}_ncu_misc.log 2>&1; echo "ncu misc rc=$?"
# source-level captures (small): the stem and the first pool -> pool_proj launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16x2 -c 1 -o ${P}_stem_src -f python bench.py $S > /dev/null 2>&1; echo "ncu stem rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_f16x2 --launch-skip 6 -c 1 -o ${P}_poolproj_src -f python bench.py $S > /dev/null 2>&1; echo "ncu poolproj rc=$?"
for k in conv pool dw misc; do python tools/ncu_summary.py full ${P}_${k}_full.ncu-rep ${P}_${k}_full.txt; done
[ $(du -sm gpurun_out | cut -f1) -gt 55 ] && rm -f ${P}_conv_full.ncu-rep
ls -la gpurun_out/r2v*; du -sh gpurun_out
