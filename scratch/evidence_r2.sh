#!/bin/bash
# Round-2 evidence run: bench lines of every BASELINE config, ncu launch lists and full-set captures.  usage: evidence_r2.sh bench|ncu
# (two gpurun calls: everything written under gpurun_out/ by one call must stay below 64 MiB)
set -x
mkdir -p gpurun_out
if [ "$1" = "bench" ]; then
python bench.py --steps 10 2>/dev/null | tail -1 > gpurun_out/r2_bench_headline.json
python bench.py --config 3 --steps 5 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_w4a8_b32.json
MQB200_W4_FUSED=0 python bench.py --config 3 --steps 5 --no-calib --no-cpu-baseline --no-decode 2>/dev/null | tail -1 > gpurun_out/r2_bench_w4a8_b32_unfused.json
python bench.py --config 5 --steps 5 --no-calib --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2_bench_gemma_s2048.json
python bench.py --config 4 --steps 5 --no-cpu-baseline --calib-samples 48 2>/dev/null | tail -1 > gpurun_out/r2_bench_stablelm.json
python scratch/calib512.py 512 2>/dev/null | tail -1
# launch list of one bench step (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile-step --no-calib --no-decode --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_calib_launches.csv python scratch/prof_calib_kernels.py > /dev/null 2>&1
python scratch/bench_decode_kernels.py 2>/dev/null | head -16 > gpurun_out/r2_decode_kernels.txt
else
# full-set captures: block 0 of the headline step (qnorm, qgemm QKV, qrope, qattn_tc, qgemm o, qnorm, qgemm w13, qgemm w2)
ncu --set full --clock-control none --profile-from-start off -k regex:"qgemm|qattn|qnorm|qrope" -c 8 -o /tmp/r2_block0 python bench.py --profile-step --no-calib --no-decode --no-cpu-baseline > /dev/null 2>&1
# W4A8 fused GEMM (config 3 shapes): the four GEMMs of block 0
ncu --set full --clock-control none --profile-from-start off -k regex:"qgemm" -c 4 -o /tmp/r2_qgemm_w4 python bench.py --config 3 --profile-step --no-calib --no-decode --no-cpu-baseline > /dev/null 2>&1
# calibration kernels: one eager step on 2 layers
ncu --set full --clock-control none -k regex:"fq_fwd|fq_bwd|wprep_rowminmax|wprep_quant|wprep_bwd_stats|wprep_bwd_apply|minmax_kernel|adam_" -s 400 -c 24 -o /tmp/r2_calib_kernels python scratch/prof_calib_kernels.py > /dev/null 2>&1
# the reports stay on the box (64 MiB limit on gpurun_out/): summarise here
python scratch/summarize_ncu.py gpurun_out/r2_ncu_block0.md "ncu full-set captures, round 2: block 0 of the headline step (batch 8 x seq 1024, TinyLlama shapes) -- qnorm, qgemm QKV (QUANT), qrope, qattn_tc, qgemm o_proj (RESID), qnorm, qgemm w1||w3 (ACTMUL), qgemm w2 (RESID, CTA pair)" /tmp/r2_block0.ncu-rep
python scratch/summarize_ncu.py gpurun_out/r2_ncu_qgemm_w4.md "ncu full-set captures, round 2: the four fused int4 x int8 GEMMs (mq_qgemm_w4a8) of block 0 at batch 32 x seq 1024" /tmp/r2_qgemm_w4.ncu-rep
python scratch/summarize_ncu.py gpurun_out/r2_ncu_calib_kernels.md "ncu full-set captures, round 2: calibration kernels of one eager e2equant training step (2 TinyLlama-shape layers, seq 1024)" /tmp/r2_calib_kernels.ncu-rep
fi
ls -la gpurun_out | tail -20; du -sh gpurun_out
