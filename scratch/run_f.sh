#!/bin/bash
# r1d: qattn single-QK-pass kernel: parity + micro-bench + step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -8
MQB200_DEBUG=1 timeout 600 python scratch/bench_qattn.py 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-calib 2>&1 | tail -1 > gpurun_out/r1d_bench.json
python -c "import json; d=json.load(open('gpurun_out/r1d_bench.json')); print(d['value'], d['ms_per_step'], d['kernel_shares'], d['roofline']['achieved'])"
