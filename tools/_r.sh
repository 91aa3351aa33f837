python -m pytest tests -x -q -m gpu > gpurun_out/r3z_tests.log 2>&1; echo rc=$? >> gpurun_out/r3z_tests.log
python bench.py --steps 100 --warmup 24 > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err
