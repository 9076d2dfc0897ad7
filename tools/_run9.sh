python -m pytest tests/test_gpu_wave_tail.py tests/test_gpu_dropin.py -q -x 2>&1 | tail -2
python bench.py --workload inference --steps 5 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_inference_1gpu_r02b.json 2> gpurun_out/bench_inf_err.log; head -c 400 gpurun_out/bench_inference_1gpu_r02b.json; tail -2 gpurun_out/bench_inf_err.log | cut -c1-200
