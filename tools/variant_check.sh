#!/bin/bash
# One GPU visit: parity tests of the main-pass kernel variants, their timings, pipeline timelines.
# usage (under gpurun): bash tools/variant_check.sh <tag>
tag=${1:-vc}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
tail -5 gpurun_out/${tag}_tests.log
VARIANTS=0,1 ITERS=24 timeout 200 python tools/dec_bench.py > gpurun_out/${tag}_dec_bench.log 2>&1
echo "dec_bench rc=$?"
grep -E '"|us' gpurun_out/${tag}_dec_bench.log | grep -E 'variant|"us"' | paste - - | head -40
for v in 0 1; do
  GLUE_TC_VARIANT=$v timeout 100 python tools/tc_trace.py > gpurun_out/${tag}_trace_v$v.log 2>&1
  echo "trace v$v rc=$?"
done
