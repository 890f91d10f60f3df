#!/bin/bash
# Round-2 evidence for the calibration path after the per-block fusions: the full 512-sample run, the launch list of an eager
# 2-layer run and full-set ncu captures of every calibration kernel.  Reports stay on the box (64 MiB limit on gpurun_out/).
set -x
mkdir -p gpurun_out
python scratch/calib512.py 512 2>/dev/null | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_calib_launches.csv python scratch/prof_calib_kernels.py > /dev/null 2>&1
# the second sample's training step: skip the FP pass and the first (warm-up) step
ncu --set full --clock-control none -k regex:"attn_probs|silu_gate|rmsnorm_l2|qkv_rope|fq_fwd|fq_bwd|wprep_rowminmax|wprep_quant|wprep_bwd_stats|wprep_bwd_apply|fold_cols|adam_" -s 120 -c 30 -o /tmp/r2_calib_kernels python scratch/prof_calib_kernels.py > /dev/null 2>&1
python scratch/summarize_ncu.py gpurun_out/r2_ncu_calib_kernels.md "ncu full-set captures, round 2: calibration kernels of one eager e2equant training step (2 TinyLlama-shape layers, seq 1024) after the per-block fusions" /tmp/r2_calib_kernels.ncu-rep
ls -la gpurun_out | tail; du -sh gpurun_out
