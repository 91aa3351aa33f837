#!/bin/bash
# Run under gpurun: plain run first (must exit 0), then the ncu launch list and one
# `--set full` capture of the fused decoder kernels (ATAC call = 2nd decoder call).
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'dec_|graphconv|graphdec' -s 6 -c 14 \
    -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
