#!/bin/bash
# up-conv (1x1, depth-to-space) shapes with forced N_TILE / CTA-group variants
for cfg in "1 128 256 16 126 126 0 0 1" "1 256 512 16 64 64 0 0 1" "1 512 1024 16 32 32 0 0 1" "1 1024 2048 16 16 16 0 0 1"; do
  for nt in 0 64 128 256; do
    for cg in 1 2; do
      echo "cfg $cfg n_tile $nt cg $cg: $(timeout 60 ./tools/probe conv $cfg $nt -1 0 $cg | grep -E 'TFLOP|FAIL|failed|error' | tr '\n' ' ' | cut -c1-110)"
    done
  done
done
NIND_TRACE=1 timeout 60 ./tools/probe conv 1 128 256 16 126 126 0 0 1
