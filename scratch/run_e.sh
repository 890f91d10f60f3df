#!/bin/bash
# round-1 evidence pack: full GPU test suite, bench lines of the BASELINE configs, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 2>/dev/null | tail -1 > gpurun_out/r1c_bench_headline.json
timeout 900 python bench.py --steps 5 --wbits 4 --batch 32 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1c_bench_w4a8_b32.json
timeout 900 python bench.py --steps 5 --model gemma-2b --seqlen 2048 --batch 4 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1c_bench_gemma_s2048.json
timeout 900 python bench.py --steps 5 --model stablelm-2-1.6b --batch 8 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1c_bench_stablelm.json
for f in gpurun_out/r1c_bench_*.json; do echo $f; python -c "import json,sys; d=json.load(open('$f')); print(d['config']['workload'], round(d['value']), 'tok/s', round(d['ms_per_step'],2), 'ms', d['roofline']['frac'], d.get('calib',{}).get('value'), d['kernel_shares'])"; done
bash scratch/prof_r1.sh r1c > /dev/null 2>&1
ls -la gpurun_out | grep r1c
