#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"qattn|qrope" -c 2 -f -o gpurun_out/r1b_attn python bench.py --profile-step > gpurun_out/r1b_attn.log 2>&1
timeout 900 python scratch/prof_calib.py 22 8 > gpurun_out/r1b_calib.log 2>&1
tail -5 gpurun_out/r1b_calib.log
