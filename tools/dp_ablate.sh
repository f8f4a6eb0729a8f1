#!/bin/bash
# dcn_pair: A/B of build variants (EDVR_B200_LIB) on one box
for rep in 1 2; do
for v in "" _cfgB _cfgE; do
  for s in 0.02 3; do
    echo -n "lib$v  "
    EDVR_B200_LIB=$PWD/edvr_b200/libedvr_b200$v.so EDVR_B200_DCN_SITE=pair timeout 120 python tools/one_site.py 28 180 320 128 8 $s 3 2>&1 | tail -1 | cut -c1-120
  done
done
done
