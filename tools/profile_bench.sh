#!/bin/bash
# Run under gpurun (1 GPU).  Every ncu pass is preceded by the same command without ncu.
#   1. launch list of a few graph-replayed training steps (shares per kernel)
#   2. `--set full` capture of the tensor-core decoder kernels at the ATAC shape
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
STEP_CMD="python bench.py --profile --steps 1 --warmup 3"
$STEP_CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
# skip the two eager warm-up steps and the capture; ~1.3 replayed steps are listed
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $STEP_CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
export ONLY=atac ITERS=2
python tools/dec_bench.py > gpurun_out/${TAG}_dec_plain.log 2>&1 || { echo "plain decoder run failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_main_kernel|tc_rowstats" -s 9 -c 3 \
    -o gpurun_out/${TAG}_decoder python tools/dec_bench.py > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
