#!/bin/bash
# round 2, GPU call A: pixel-pair mode first light (prebuilt tools/probe_pm) + ncu of the first layer
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
for cfg in "9 64 64 2 18 34 0 0 0" "9 64 64 4 250 250 0 0 0" "9 64 64 4 254 254 0 0 2" "9 128 64 4 252 252 0 0 0" "9 64 64 32 250 250 0 0 0" "9 128 64 32 252 252 0 0 0" "9 64 64 32 254 254 0 0 2"; do
  for pair in 0 1; do
    echo "$cfg pair $pair: $(timeout 60 ./tools/probe_pm conv $cfg 0 -1 0 0 -1 $pair 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch' | head -6 | tr '\n' ' ' | sed 's/checked=[0-9]* //' | cut -c1-400)"
  done
done
echo "--- first layer (c8) probe"
timeout 60 ./tools/probe conv 9 8 64 32 252 252 0 0 0 | grep -E "CONV|TFLOP|PASS|FAIL"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/r02_c8 ./tools/probe conv 9 8 64 32 252 252 0 0 0 > gpurun_out/r02_c8_ncu.log 2>&1
echo "ncu c8 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/r02_pm64 ./tools/probe_pm conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 > gpurun_out/r02_pm64_ncu.log 2>&1
echo "ncu pm64 exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/r02_pm128 ./tools/probe_pm conv 9 128 64 32 252 252 0 0 0 0 -1 0 0 -1 1 > gpurun_out/r02_pm128_ncu.log 2>&1
echo "ncu pm128 exit $?"
