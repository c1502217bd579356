#!/bin/bash
# ncu launch list for the bench command (B200_PROFILING.md recipe): plain run first, then the two timed steps
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; exit 1; }
cut -c1-400 gpurun_out/ncu_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 279 -c 186 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
