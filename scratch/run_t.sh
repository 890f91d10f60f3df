#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/r1d_bench_2gpu.json
python -c "import json; d=json.load(open('gpurun_out/r1d_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -x -q 2>&1 | tail -4
