#!/bin/bash
# ncu --set full of the 64->64 3x3 layer shape through the probe (plain run first)
CFG="conv 9 64 64 8 254 254 0 0 0"
./tools/probe $CFG > gpurun_out/probe64_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
grep -E "TFLOP|PASS|FAIL" gpurun_out/probe64_plain.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:igemm -c 1 -f -o gpurun_out/probe64 ./tools/probe $CFG > gpurun_out/probe64_ncu.log 2>&1
echo "ncu exit $?"
