python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "fused_encoder or act_bn" > gpurun_out/r3j_tests.log 2>&1; echo rc=$? >> gpurun_out/r3j_tests.log
python tools/mlp_bench.py > gpurun_out/r3j_mlp.json 2> gpurun_out/r3j_err.txt
