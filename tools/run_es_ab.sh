#!/bin/bash
# A/B of the number of epilogue sets for N_TILE=64 kernels on the same box
for rep in 1 2; do
for bin in probe_es2 probe probe_es4; do
  for cfg in "9 64 64 1 510 510 0 0 0" "9 64 64 1 510 510 0 0 2" "9 128 64 1 508 508 0 0 0" "1 64 64 1 506 506 0 0 0"; do
    echo "$bin $cfg: $(timeout 60 ./tools/$bin conv $cfg | grep -E 'TFLOP|FAIL' | tr '\n' ' ')"
  done
done
done
