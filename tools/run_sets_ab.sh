#!/bin/bash
for rep in 1 2; do
for bin in probe probe_x16 probe_x16s3; do
  for cfg in "9 64 64 8 254 254 0 0 2" "9 64 64 8 254 254 0 0 0" "9 128 64 8 252 252 0 0 0" "9 8 64 8 252 252 0 0 0" "9 64 64 8 254 254 0 0 0 0 -1 0 1"; do
    echo "$bin $cfg: $(timeout 60 ./tools/$bin conv $cfg | grep -E 'TFLOP|FAIL|PASS|failed' | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-150)"
  done
done
done
