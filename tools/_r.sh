python -m pytest tests -x -q -m gpu > gpurun_out/r4h_tests.log 2>&1; echo rc=$? >> gpurun_out/r4h_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4h_smoke.log 2>&1
python bench.py --steps 100 --warmup 24 > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err
