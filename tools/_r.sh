python -m pytest tests/test_step_gpu.py -x -q -m gpu > gpurun_out/r4k_tests.log 2>&1; echo rc=$? >> gpurun_out/r4k_tests.log
python bench.py --steps 60 --warmup 16 --skip-cpu-baseline --no-fit > gpurun_out/r4k_bench.json 2> gpurun_out/r4k_bench.err
python tools/step_profile.py > gpurun_out/r4k_timeline.txt 2>&1
