for c in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 40 --warmup 12 --skip-cpu-baseline --no-fit > gpurun_out/r4c_bench_$c.json 2> gpurun_out/r4c_bench_$c.err
done
bash tools/profile_bench.sh r4c > gpurun_out/r4c_profile.log 2>&1
python tools/step_profile.py > gpurun_out/r4c_timeline.txt 2>&1
python tools/gnll_bench.py > gpurun_out/r4c_gnll.json 2>/dev/null
python tools/mlp_bench.py > gpurun_out/r4c_mlp.json 2>/dev/null
