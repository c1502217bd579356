#!/bin/bash
# round 2, GPU call J: software-pipelined TMEM loads in the epilogue of the 384-thread kernels (next 32-column chunk in
# flight while the current one is processed).  Same-box A/B against the previous commit (tools/build_base.sh).
run() { P=$1; shift; echo "--- $(basename $P) $*: $(timeout 60 $P "$@" 2>&1 | grep -E 'TFLOP|FAIL|PASS|failed|error|mismatch|pool:|stray' | head -8 | tr '\n' ' ' | sed 's/checked=[0-9]* //; s/maxerr.*bad=/bad=/' | cut -c1-200)"; }
for P in ./tools/probe_base ./tools/probe; do
run $P conv 1 128 256 133 126 126 0 0 1
run $P conv 1 256 512 133 62 62 0 0 1
run $P conv 1 512 1024 133 30 30 0 0 1
run $P conv 1 1024 2048 133 14 14 0 0 1
run $P conv 9 64 128 32 126 126 0 0 0
run $P conv 9 128 128 32 124 124 0 0 0
run $P conv 9 256 128 32 128 128 0 0 0
run $P conv 9 256 256 32 60 60 0 0 0
run $P conv 9 512 512 32 30 30 0 0 0
done
echo "=== pytest (kernel tests)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "each_layer or variants or forward or unet or golden" 2>&1 | tail -4
echo "=== A/B bench"
bash tools/run_ab.sh 3
