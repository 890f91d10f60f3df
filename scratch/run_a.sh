#!/bin/bash
# tests + kernel micro-bench + bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scratch/bench_qgemm.py 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --no-calib 2>&1 | tail -3
