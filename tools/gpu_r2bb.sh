#!/bin/bash
# GPU session BB: ncu evidence of the committed end-of-round build (launch lists + DRAM traffic of GoogLeNet / SSD / mnist_bn,
# --set full of the 40 contraction launches of one GoogLeNet pass and of the SSD depthwise layers)
mkdir -p gpurun_out
P=gpurun_out/r2bb
B="--no-secondary --no-f16 --sustain 0 --cpu-budget 1"
S="--steps 2 --warmup 3 $B"
python bench.py $S > /dev/null 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file ${P}_launches_googlenet.csv python bench.py $S > ${P}_ncu_g.log 2>&1; echo "ncu launches(g) rc=$?"
python bench.py $S --workload ssd_mobilenet_v1_coco > /dev/null 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file ${P}_launches_ssd.csv python bench.py $S --workload ssd_mobilenet_v1_coco > ${P}_ncu_s.log 2>&1; echo "ncu launches(ssd) rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv_f16x2 -c 40 -o ${P}_conv_full -f python bench.py $S > ${P}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
python tools/ncu_summary.py full ${P}_conv_full.ncu-rep ${P}_conv_full.txt
rm -f ${P}_conv_full.ncu-rep
ls -la gpurun_out
