python -m pytest tests/test_gpu_ops.py -q -x -k "ws_stacked or bn_stats_from_conv or conv_tc_bf16" 2>&1 | tail -30 > gpurun_out/t3.txt
python tools/conv_shapes.py --only wgrad --reps 5 --sweep wgws_mstack=0,1 > gpurun_out/shapes3c.txt 2>&1
tail -30 gpurun_out/t3.txt
