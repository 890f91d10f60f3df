#!/bin/bash
echo "--- normal"; timeout 300 python scratch/bench_qgemm.py 1 2 2>&1 | grep -v "M=32768" | tail -12
echo "--- epilogue skipped"; MQ_QGEMM_DBG=1 timeout 300 python scratch/bench_qgemm.py 1 2 4 2>&1 | grep -v "M=32768" | tail -19
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
