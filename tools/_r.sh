python -m pytest tests/test_step_gpu.py -x -q -m gpu > gpurun_out/r3y_step.log 2>&1; echo rc=$? >> gpurun_out/r3y_step.log
python bench.py --steps 60 --warmup 16 > gpurun_out/r3y_bench.json 2> gpurun_out/r3y_bench.err
python tools/step_profile.py > gpurun_out/r3y_timeline.txt 2>&1
