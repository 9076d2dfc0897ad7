python -m pytest tests/test_gpu_step.py -q -x 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_1gpu_r02h.json 2> gpurun_out/bench_err.log; head -c 300 gpurun_out/bench_1gpu_r02h.json; tail -3 gpurun_out/bench_err.log | cut -c1-200
echo; echo "== multi-graph"; python bench.py --steps 5 --warmup 3 --no-gpu-eager --no-cpu-baseline --multi-graph 2>&1 | tail -1 | cut -c1-220
echo "== overlap off"; VCA_OVERLAP_OPT=0 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
