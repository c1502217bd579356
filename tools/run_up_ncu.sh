#!/bin/bash
for epi in 0 1; do
  echo "epi $epi: $(timeout 60 ./tools/probe conv 1 128 256 16 126 126 0 0 $epi | grep -E 'TFLOP|FAIL' | tr '\n' ' ')"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -c 2 -o gpurun_out/up4_probe -f ./tools/probe conv 1 128 256 16 126 126 0 0 1 > gpurun_out/up4_ncu.log 2>&1
tail -3 gpurun_out/up4_ncu.log
