python bench.py > gpurun_out/bench_1gpu_r02_final.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_1gpu_r02_final.json; tail -3 gpurun_out/bench_err.log | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_cpu_r02.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_reference_cpu_r02.json | cut -c1-600
