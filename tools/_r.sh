python -m pytest tests/test_step_gpu.py -x -q -m gpu > gpurun_out/r4a_step.log 2>&1; echo rc=$? >> gpurun_out/r4a_step.log
python bench.py --steps 100 --warmup 24 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err
