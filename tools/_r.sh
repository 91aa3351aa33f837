python -m pytest tests/test_step_gpu.py tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/r3r_tests.log 2>&1; echo rc=$? >> gpurun_out/r3r_tests.log
python bench.py --steps 60 --warmup 16 > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err
python tools/step_profile.py > gpurun_out/r3r_timeline.txt 2>&1
