#!/bin/bash
# round-1 session-d evidence pack: full GPU test suite, bench lines of the BASELINE configs, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r1d_pytest_gpu.txt
timeout 900 python bench.py --steps 10 --decode-batch 32 2>/dev/null | tail -1 > gpurun_out/r1d_bench_headline.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1d_bench_reference_arm.json
timeout 900 python bench.py --steps 5 --wbits 4 --batch 32 --no-calib --no-cpu-baseline --no-decode 2>/dev/null | tail -1 > gpurun_out/r1d_bench_w4a8_b32.json
timeout 900 python bench.py --steps 5 --model gemma-2b --seqlen 2048 --batch 8 --decode-batch 32 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r1d_bench_gemma_s2048.json
timeout 900 python bench.py --steps 5 --model stablelm-2-1.6b --batch 8 --no-calib --no-cpu-baseline --no-decode 2>/dev/null | tail -1 > gpurun_out/r1d_bench_stablelm.json
for f in gpurun_out/r1d_bench_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(d.get('config',{}).get('workload'), round(d.get('value',0)), d.get('unit'), round(d.get('ms_per_step',0),2), 'ms', (d.get('roofline') or {}).get('frac'), (d.get('calib') or {}).get('value'), (d.get('decode') or {}).get('value'), d.get('kernel_shares'))"; done
CMD="python bench.py --profile-step --batch 8"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1d_launches.csv $CMD > gpurun_out/r1d_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:qgemm -c 4 -f -o gpurun_out/r1d_qgemm $CMD > gpurun_out/r1d_qgemm.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"qattn|qrope|qnorm" -c 3 -f -o gpurun_out/r1d_attn $CMD > gpurun_out/r1d_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"qgemv|qattn_decode|qnorm_row" -c 12 -f -o gpurun_out/r1d_decode python scratch/prof_decode.py tinyllama-1.1b 1024 8 > gpurun_out/r1d_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fgemv -c 1 -f -o gpurun_out/r1d_fgemv python scratch/prof_decode.py tinyllama-1.1b 1024 8 > gpurun_out/r1d_fgemv.log 2>&1
ls -la gpurun_out | grep r1d
