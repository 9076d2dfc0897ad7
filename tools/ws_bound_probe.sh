#!/bin/bash
# What bounds the weights-stationary conv kernel?  The same launches with parts of the kernel switched off ("ws_dbg":
# 1 no global stores, 2 no TMEM reads / stores, 4 no MMAs, 8 / 16 MMA N forced to 128 / 256 -- results are garbage, only
# the time matters).   bash tools/ws_bound_probe.sh > gpurun_out/ws_bound_probe.txt
for f in resnet.layer1 gen.g2.1 gen.g3.pair v_front.stem; do
  python tools/conv_shapes.py --filter "$f" --only fwd --reps 5 --sweep ws_dbg=0,1,2,4,6,8,16,10,18 2>&1 | grep -v Warn
done
python tools/conv_shapes.py --filter "dis." --reps 5 --sweep ws_min_taps=2,1 2>&1 | grep -v Warn
