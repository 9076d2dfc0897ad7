python -m pytest tests/test_gpu_ops.py -q -x -k "maxpool or fresh_autograd" 2>&1 | tail -3
python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_targets_plain.log; exit 1; }
VCA_NCU=1 timeout 1200 ncu --set full --clock-control none -k regex:'conv_tc|att_fwd|bmm_tc|bn_prelu|gl_frames|gl_ola|unslab' --launch-count 40 -f -o /tmp/r02_kernels python tools/ncu_targets.py > gpurun_out/ncu_targets_ncu.log 2>&1
tail -3 gpurun_out/ncu_targets_ncu.log; ls -la /tmp/r02_kernels.ncu-rep
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/r02_kernels_raw.csv
sz=$(stat -c %s /tmp/r02_kernels.ncu-rep); if [ "$sz" -lt 45000000 ]; then cp /tmp/r02_kernels.ncu-rep gpurun_out/; fi
python tools/step_trace.py 32 75 r02b > gpurun_out/step_trace_r02b.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline > gpurun_out/bench_1gpu_r02g.json 2> gpurun_out/bench_err.log; head -c 300 gpurun_out/bench_1gpu_r02g.json
ls -la gpurun_out | head -20
