#!/bin/bash
for b in 8 32 64; do
python bench.py --steps 3 --no-cpu-baseline --no-calib --decode-batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['decode']; print('tinyllama', d['batch'], round(d['value']), 'tok/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3))"
done
for b in 8 32; do
python bench.py --steps 2 --model gemma-2b --seqlen 2048 --batch 4 --no-cpu-baseline --no-calib --decode-batch $b 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['decode']; print('gemma', d['batch'], round(d['value']), 'tok/s', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],3))"
done
