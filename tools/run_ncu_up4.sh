#!/bin/bash
# source-level ncu of the up4 shape (1x1, 128 -> 4x64 depth-to-space, 133 crops): which instructions generate the
# shared-memory wavefronts of this epilogue-bound kernel
./tools/probe conv 1 128 256 133 126 126 0 0 1 | grep -E "TFLOP|PASS"
ncu --set full --import-source on --clock-control none -k regex:igemm -s 1 -c 1 -f -o gpurun_out/up4_src ./tools/probe conv 1 128 256 133 126 126 0 0 1 > gpurun_out/up4_src.log 2>&1
ncu -i gpurun_out/up4_src.ncu-rep --page source --csv > gpurun_out/up4_src.csv 2>/dev/null
ncu -i gpurun_out/up4_src.ncu-rep --page raw --csv > gpurun_out/up4_raw.csv 2>/dev/null
rm -f gpurun_out/up4_src.ncu-rep
ls -la gpurun_out/up4_src.csv
