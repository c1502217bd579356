#!/bin/bash
# round 2, GPU call G: warp index made provably warp-uniform (__shfl_sync broadcast) so that tile geometry and the
# TMA-store coordinates live in uniform registers (no R2UR waterfall loops in the epilogue): same-box A/B against the
# previous kernel (tools/probe_base, built from the parent commit's igemm.cuh), then the bench with the layer table.
run() { P=$1; shift; echo "--- $(basename $P) $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
for P in ./tools/probe_base ./tools/probe; do
run $P conv 9 8 64 32 252 252 0 0 0
run $P conv 1 128 256 32 126 126 0 0 1
run $P conv 1 256 512 32 62 62 0 0 1
run $P conv 9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 1
run $P conv 9 64 64 32 254 254 0 0 2 0 -1 0 0 -1 1 0
run $P conv 9 128 64 32 252 252 0 0 0
run $P conv 9 64 128 32 126 126 0 0 0
run $P conv 9 128 128 32 124 124 0 0 0
run $P conv 9 256 128 32 128 128 0 0 0
run $P conv 9 256 256 32 60 60 0 0 0
run $P conv 9 512 512 32 30 30 0 0 0
done
echo "=== bench"
python bench.py --steps 5 --no-cpu-baseline --layers 2> gpurun_out/r2g_layers.txt > gpurun_out/r2g_bench.json; cut -c1-200 gpurun_out/r2g_bench.json
grep -o '"parity".\{0,700\}' gpurun_out/r2g_bench.json | cut -c1-700
