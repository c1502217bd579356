#!/bin/bash
# EXPERIMENTAL pixel-pair mode: first hardware check for the next round.  Build the probe with the
# mode compiled in, then for each C_out = 64 layer shape compare pair=0 / pair=1 against the naive convolution.
# args: taps cin n B Hs Ws 0 0 epi n_tile ws ctas cg flat pair
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DNIND_PAIR_MODE=1 -o tools/probe_pm tools/probe.cu || exit 1
for cfg in "9 64 64 2 18 34 0 0 0" "9 64 64 4 250 250 0 0 0" "9 64 64 4 254 254 0 0 2" "9 128 64 4 252 252 0 0 0"; do
  for pair in 0 1; do
    echo "$cfg pair $pair: $(timeout 60 ./tools/probe_pm conv $cfg 0 -1 0 0 -1 $pair 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error' | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-130)"
  done
done
