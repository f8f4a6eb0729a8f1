#!/bin/bash
# dcn_pair: timing at two offset scales, ablations, parity (one gpurun call)
for s in 0.02 3; do
  EDVR_B200_DCN_SITE=pair timeout 120 python tools/one_site.py 28 180 320 128 8 $s 3 2>&1 | tail -1 | cut -c1-120
done
EDVR_B200_DCN_SITE=pair timeout 120 python tools/one_site.py 28 180 320 128 8 0.02 3 2>&1 | tail -1 | cut -c1-120
EDVR_B200_DCN_SITE=fused timeout 120 python tools/one_site.py 28 180 320 128 8 0.02 3 2>&1 | tail -1 | cut -c1-120
EDVR_B200_DCN_SITE=pair timeout 120 python tools/one_site.py 16 64 64 64 8 0.02 3 2>&1 | tail -1 | cut -c1-120
EDVR_B200_DCN_SITE=fused timeout 120 python tools/one_site.py 16 64 64 64 8 0.02 3 2>&1 | tail -1 | cut -c1-120
timeout 200 python tools/dcn_sweep.py --modes pair --n 4 --sigmas 0.02,3,10 | grep parity | cut -c1-220
