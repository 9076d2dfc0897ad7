python -m pytest tests/test_gpu_ops.py -q -x -k "ws_stacked or bn_stats_from_conv or conv_tc_bf16" 2>&1 | tail -30 > gpurun_out/t3.txt
python tools/conv_shapes.py --only fwd --reps 5 > gpurun_out/shapes3a.txt 2>&1
python tools/conv_shapes.py --only fwd --reps 5 --filter "resnet.layer1" --sweep ws_dbg=0,1,2,4,6 >> gpurun_out/shapes3a.txt 2>&1
python tools/conv_shapes.py --reps 5 --filter "resnet.layer2" --ws 2 --opt ws_wbudget_kb=148 --tag "force148 " >> gpurun_out/shapes3a.txt 2>&1
python tools/conv_shapes.py --reps 5 --filter "resnet.layer2" --ws 2 --opt ws_wbudget_kb=148 --opt ws_tma_out=0 --tag "force148 notma " >> gpurun_out/shapes3a.txt 2>&1
tail -30 gpurun_out/t3.txt
