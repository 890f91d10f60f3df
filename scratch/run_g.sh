#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:qattn4 -c 1 -s 1 -f -o gpurun_out/r1d_qattn4 python scratch/prof_qattn.py > gpurun_out/r1d_qattn4.log 2>&1
tail -3 gpurun_out/r1d_qattn4.log
ls -la gpurun_out
