#!/bin/bash
# one --set full capture of conv launches of the first timed batch (after the same command ran clean):
# 3 warm-up steps x 4 batches x 22 conv launches = 264; +17 = tconvs3.0, tconvs3.2, up4, tconvs4.0, tconvs4.2+head of the first timed batch
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain2.json 2> gpurun_out/ncu_plain2.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:igemm -s 281 -c 5 -f -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
