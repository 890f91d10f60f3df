#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r1d_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r1d_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --decode-batch 32 2>/dev/null | tail -1 > gpurun_out/r1d_bench_headline.json
python -c "
import json
d=json.load(open('gpurun_out/r1d_bench_headline.json'))
print(round(d['value']), 'tok/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'calib', d['calib']['value'], 'decode', d['decode']['value'], d['kernel_shares']['qattn'])"
CMD="python bench.py --profile-step --model gemma-2b --seqlen 2048 --batch 8"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1d_launches_gemma.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:qattn4 -c 1 -f -o gpurun_out/r1d_attn_gemma $CMD > /dev/null 2>&1
ls -la gpurun_out | grep -E "gemma|headline"
