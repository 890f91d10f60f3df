#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --no-cpu-baseline --no-calib 2>&1 | tail -1 > gpurun_out/r1d_bench_tl.json
python -c "import json; d=json.load(open('gpurun_out/r1d_bench_tl.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], json.dumps(d['decode']))"
timeout 900 python bench.py --steps 3 --model gemma-2b --seqlen 2048 --batch 8 --no-calib --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1d_bench_gemma.json
python -c "import json; d=json.load(open('gpurun_out/r1d_bench_gemma.json')); print(d['value'], d['ms_per_step'], d['kernel_shares'], json.dumps(d['decode']))"
