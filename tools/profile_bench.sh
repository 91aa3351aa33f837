#!/bin/bash
# Run under gpurun (1 GPU).  Every ncu pass is preceded by the same command without ncu.
#   1. launch list of a few graph-replayed training steps (shares per kernel)
#   2. DRAM traffic of the fused decoder call of the ATAC modality (three loss terms)
#   3. `--set full` capture of its two main kernels
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
STEP_CMD="python bench.py --profile --steps 3 --warmup 3"
$STEP_CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain step run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $STEP_CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv "one replayed training step (cfg2, between two input loads)" 0 multi_copy_kernel \
    > gpurun_out/${TAG}_launches_summary.md 2>&1
python tools/ncu_decoder.py > gpurun_out/${TAG}_dec_plain.log 2>&1 || { echo "plain decoder run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_dec_traffic.csv python tools/ncu_decoder.py > gpurun_out/${TAG}_ncu2.log 2>&1
echo "traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:main3_kernel -s 6 -c 2 \
    -o gpurun_out/${TAG}_dec_main3 python tools/ncu_decoder.py > gpurun_out/${TAG}_ncu3.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
