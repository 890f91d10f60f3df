#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_qgemm_gpu.py -m gpu -x -q -k "cluster_variants" 2>&1 | tail -8
timeout 300 python scratch/bench_qgemm.py 2>&1 | tail -36
