python -m pytest tests/test_gpu_step.py -q -x -s -k "reproducible" 2>&1 | grep -E "run-to-run|passed|failed|Error|assert" | head
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/gpu_tests_r02d.txt; tail -3 gpurun_out/gpu_tests_r02d.txt
python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_1gpu_r02j.json 2> gpurun_out/bench_err.log; head -c 300 gpurun_out/bench_1gpu_r02j.json
