#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -2
