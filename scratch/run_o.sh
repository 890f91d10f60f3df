#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1d_decode_launches.csv python scratch/prof_decode.py tinyllama-1.1b 1024 8 > gpurun_out/r1d_decode_launches.log 2>&1
tail -2 gpurun_out/r1d_decode_launches.log
