#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  $1 = tag
TAG=${1:-r1}
mkdir -p gpurun_out
CMD="python bench.py --profile-step --batch 8"
# (1) launch list of one step (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
# (2) full captures of the top kernels: first decoder block's four GEMMs, attention, rope, norm
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:qgemm -c 4 \
    -f -o gpurun_out/${TAG}_qgemm $CMD > gpurun_out/${TAG}_qgemm.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"qattn|qrope|qnorm" -c 3 \
    -f -o gpurun_out/${TAG}_attn $CMD > gpurun_out/${TAG}_attn.log 2>&1
ls -la gpurun_out/
