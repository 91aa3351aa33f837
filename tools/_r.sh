python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "fused_encoder or act_bn" > gpurun_out/r3o_tests.log 2>&1; echo rc=$? >> gpurun_out/r3o_tests.log
python tools/mlp_bench.py > gpurun_out/r3o_mlp.json 2> gpurun_out/r3o_err.txt
python bench.py --steps 60 --warmup 16 > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err
python tools/step_profile.py > gpurun_out/r3o_timeline.txt 2>&1
