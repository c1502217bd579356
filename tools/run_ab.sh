#!/bin/bash
# Same-box A/B of the whole bench step: libnind_b200_base.so (tools/build_base.sh) vs the working tree, alternating.
# usage: run_ab.sh [rounds] [extra bench args]
R=${1:-2}; shift
for i in $(seq 1 $R); do
for which in base new; do
  if [ $which = base ]; then export NIND_LIB=$PWD/nind_denoise_b200/libnind_b200_base.so; else unset NIND_LIB; fi
  python bench.py --steps 6 --no-cpu-baseline --no-parity --images 2 --layers "$@" 2> gpurun_out/ab_${which}_$i.layers > gpurun_out/ab_${which}_$i.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${which}_$i.json"))
r=d["roofline"]
print("$which $i: %.1f MP/s  %.2f ms/step  e2e %.1f  conv kernels %.2f ms  outside %.2f ms  clk %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel_ms_per_step"], r["ms_outside_conv_kernels"], d["clocks"]["sm_mhz"]))
PY
done
done
