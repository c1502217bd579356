#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch list
# of the two timed steps and one --set full capture of four conv launches of the first timed batch.
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; exit 1; }
cat gpurun_out/ncu_plain.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 279 -c 186 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:igemm -s 281 -c 4 -f -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
