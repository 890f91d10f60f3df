#!/bin/bash
# Round-2 final evidence: GPU test log, bench lines of every BASELINE config, launch list + full-set captures of block 0.
# Everything written under gpurun_out/ by one call stays far below the 64 MiB limit (ncu reports go to /tmp).
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rs 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu.txt
python bench.py --steps 10 2>/dev/null | tail -1 > gpurun_out/r2_bench_headline.json
python bench.py --config 3 --steps 5 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_w4a8_b32.json
python bench.py --config 5 --steps 5 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_gemma_s2048.json
python bench.py --config 4 --steps 5 --no-cpu-baseline --calib-samples 144 2>/dev/null | tail -1 > gpurun_out/r2_bench_stablelm.json
for m in d s p; do MQB200_QNORM=$m python bench.py --no-calib --no-cpu-baseline --no-decode --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MQB200_QNORM=$m', d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernel_shares'].items()})"; done > gpurun_out/r2_qnorm_variants.txt
python scratch/bench_decode_kernels.py 2>/dev/null | head -24 > gpurun_out/r2_decode_kernels.txt
for p in 1 0; do MQB200_PDL=$p python bench.py --no-calib --no-cpu-baseline --steps 3 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('MQB200_PDL=$p decode', d['decode']['value'], 'tok/s', d['decode']['ms_per_step'], 'ms/step')"; done >> gpurun_out/r2_decode_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile-step --no-calib --no-decode --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --profile-from-start off -k regex:"qgemm|qattn|qnorm|qrope" -c 8 -o /tmp/r2_block0 python bench.py --profile-step --no-calib --no-decode --no-cpu-baseline > /dev/null 2>&1
python scratch/summarize_ncu.py gpurun_out/r2_ncu_block0.md "ncu full-set captures, round 2: block 0 of the headline step (batch 8 x seq 1024, TinyLlama shapes) -- qnorm, qgemm QKV (QUANT), qrope, qattn_tc, qgemm o_proj (RESID), qnorm, qgemm w1||w3 (ACTMUL), qgemm w2 (RESID, CTA pair)" /tmp/r2_block0.ncu-rep
# calibration kernels, forward and backward of the second sample's training step (eager, 2 layers)
ncu --set full --clock-control none -k regex:"attn_probs|silu_gate|rmsnorm_l2|qkv_rope|fq_bwd|fold_cols" -s 24 -c 24 -o /tmp/r2_calib_fused python scratch/prof_calib_kernels.py > /dev/null 2>&1
python scratch/summarize_ncu.py gpurun_out/r2_ncu_calib_fused.md "ncu full-set captures, round 2: the fused per-block calibration kernels (forward and backward) in an eager e2equant run on 2 TinyLlama-shape layers, seq 1024" /tmp/r2_calib_fused.ncu-rep
ls -la gpurun_out | tail -20; du -sh gpurun_out
