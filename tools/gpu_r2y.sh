#!/bin/bash
# GPU session Y: first-vs-second replay of freshly loaded networks on poisoned memory
mkdir -p gpurun_out
P=gpurun_out/r2y
timeout 300 python tools/find_race_fresh.py --trials 10 --poison nan --reuse 0 > ${P}_nan_noreuse.txt 2>&1; tail -12 ${P}_nan_noreuse.txt
timeout 300 python tools/find_race_fresh.py --trials 10 --poison nan --reuse 1 > ${P}_nan_reuse.txt 2>&1; tail -12 ${P}_nan_reuse.txt
timeout 200 python tools/find_race_fresh.py --trials 6 --poison nan --reuse 1 --workload ssd_mobilenet_v1_coco --batch 64 > ${P}_nan_ssd.txt 2>&1; tail -8 ${P}_nan_ssd.txt
