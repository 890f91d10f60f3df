#!/bin/bash
timeout 600 python -m pytest tests/test_qgemm_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scratch/bench_qgemm.py 2>&1 | tail -32
