python -m pytest tests/test_gpu_ops.py -q -x -k "tap_major" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_1gpu_r02f.json 2> gpurun_out/bench_err.log; head -c 400 gpurun_out/bench_1gpu_r02f.json
