#!/bin/bash
# same-box A/B of two probe builds (tools/probe = baseline, tools/probe_epi = candidate), PASS/FAIL + time
for bin in probe probe_epi; do
  for cfg in "9 64 64 4 510 510 0 0 0" "9 64 64 4 510 510 0 0 2" "9 128 64 4 508 508 0 0 0" "9 64 128 8 252 252 0 0 0" "9 128 128 8 252 252 0 0 0" "9 256 128 8 252 252 0 0 0" "9 512 256 8 124 124 0 0 0" "9 1024 512 16 60 60 0 0 0" "1 1024 2048 16 28 28 0 0 1" "1 512 1024 16 32 32 0 0 1" "1 256 512 16 64 64 0 0 1" "1 128 256 16 126 126 0 0 1" "9 8 64 4 508 508 0 0 0" ; do
    echo "$bin $cfg: $(timeout 60 ./tools/$bin conv $cfg | grep -E 'TFLOP|FAIL|PASS|failed' | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-150)"
  done
done
