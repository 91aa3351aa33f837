#!/bin/bash
# usage (under gpurun): bash tools/pdl_check.sh <tag>
tag=${1:-pdl}
mkdir -p gpurun_out
GLUE_TC_PDL=1 timeout 400 python -m pytest tests/test_tc_gpu.py tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
echo "tests(pdl=1) rc=$?"; tail -3 gpurun_out/${tag}_tests.log
VARIANTS=GLUE_TC_PDL=0,GLUE_TC_PDL=1 ITERS=24 timeout 200 python tools/dec_bench.py > gpurun_out/${tag}_dec_bench.log 2>&1
echo "dec_bench rc=$?"
grep -E '"|us' gpurun_out/${tag}_dec_bench.log | grep -E 'variant|"us"' | paste - - | head -40
GLUE_TC_PDL=1 timeout 300 python -m pytest tests/test_step_gpu.py -x -q -m gpu > gpurun_out/${tag}_step_tests.log 2>&1
echo "step tests(pdl=1) rc=$?"; tail -3 gpurun_out/${tag}_step_tests.log
GLUE_TC_PDL=1 timeout 300 python bench.py --skip-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench(pdl=1) rc=$?"; cut -c1-150 gpurun_out/${tag}_bench.json; grep -o '"e2e": {[^}]*}[^}]*}' gpurun_out/${tag}_bench.json | cut -c1-200
