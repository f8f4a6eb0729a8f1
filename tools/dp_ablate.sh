#!/bin/bash
# L2 set-aside for the trunk's fp32 residual stream: engine timing (cfg 3, B=4 and B=1) per set-aside size
for mb in 0 32 48 64 80 96; do
  echo -n "L2_PERSIST_MB=$mb  "
  EDVR_B200_L2_PERSIST_MB=$mb timeout 200 python tools/time_engine.py --batches 4,1 --iters 20 2>&1 | tail -3 | tr '\n' ' ' | cut -c1-400; echo
done
