python -m pytest tests/test_ops_gpu.py tests/test_step_gpu.py -x -q -m gpu -k "graph_encoder or step or fit" > gpurun_out/r4i_tests.log 2>&1; echo rc=$? >> gpurun_out/r4i_tests.log
python bench.py --steps 60 --warmup 16 --skip-cpu-baseline --no-fit > gpurun_out/r4i_bench.json 2> gpurun_out/r4i_bench.err
python tools/step_profile.py > gpurun_out/r4i_timeline.txt 2>&1
