#!/bin/bash
# Round-2 evidence for the calibration path after the per-block fusions: the full 512-sample run and the launch list of an eager
# 2-layer run (full-set ncu captures of the fused kernels: scratch/evidence_r2_final.sh).
set -x
mkdir -p gpurun_out
python scratch/calib512.py 512 2>/dev/null | tail -1
cp gpurun_out/r2_calib512.json gpurun_out/r2_calib512_final.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_calib_launches.csv python scratch/prof_calib_kernels.py > /dev/null 2>&1
python scratch/prof_calib.py 22 8 2>&1 | grep -A45 "GPU kernel time total" > gpurun_out/r2_calib_step_profile.txt
ls -la gpurun_out | tail; du -sh gpurun_out
