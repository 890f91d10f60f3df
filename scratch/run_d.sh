#!/bin/bash
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-calib 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_shares'])"
timeout 1000 python scratch/time_calib.py e2e 2>&1 | tail -5
