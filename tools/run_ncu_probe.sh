#!/bin/bash
# ncu --set full captures of three representative conv shapes through tools/probe.
i=0
for cfg in "9 64 64 1 510 510 0 0 0" "9 256 128 2 252 252 0 0 0" "9 1024 512 8 60 60 0 0 0"; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s 1 -c 1 \
      -f -o gpurun_out/prof_conv$i ./tools/probe conv $cfg > gpurun_out/ncu_conv$i.log 2>&1
  echo "cfg $cfg -> exit $?"
done
