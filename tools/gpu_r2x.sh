#!/bin/bash
# GPU session X: locate a rare replay-to-replay difference of GoogLeNet batch 256
mkdir -p gpurun_out
P=gpurun_out/r2x
python tools/determinism.py --batch 256 --iters 300 > ${P}_det_default.txt 2>&1; tail -1 ${P}_det_default.txt
python tools/find_race.py --iters 150 > ${P}_race_default.txt 2>&1; tail -6 ${P}_race_default.txt
B200OV_F16_STAGE=0 python tools/find_race.py --iters 150 > ${P}_race_nostage.txt 2>&1; tail -4 ${P}_race_nostage.txt
B200OV_F16_STEM_GROUP=1 python tools/find_race.py --iters 150 > ${P}_race_nogroup.txt 2>&1; tail -4 ${P}_race_nogroup.txt
B200OV_NO_POOL_FUSE=1 python tools/find_race.py --iters 150 > ${P}_race_nopool.txt 2>&1; tail -4 ${P}_race_nopool.txt
