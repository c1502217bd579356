#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe), round 2: the plain run first, then (1) the launch
# list of the two timed steps and (2) one --set full capture of every conv launch of the first timed batch
# (22 launches: the 23-launch plan minus the gather), converted to CSV on the box (the .ncu-rep is too big to
# bring back).  Then: python tools/summarize_ncu.py r02
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --images 2"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
cut -c1-300 gpurun_out/ncu_plain.json
# bench.py brackets its timed region with cudaProfilerStart/Stop: 2 timed steps = 2 x (4 forwards x 23 launches + 1 stitch) = 186 launches
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# igemm launches only: the first 22 of the timed region = the first forward (133 crops)
ncu --set full --clock-control none --profile-from-start off -k regex:igemm -c 22 -f -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/prof_bench.ncu-rep --page raw --csv > gpurun_out/prof_bench_raw.csv 2>/dev/null
ls -la gpurun_out/prof_bench.ncu-rep gpurun_out/prof_bench_raw.csv
rm -f gpurun_out/prof_bench.ncu-rep
