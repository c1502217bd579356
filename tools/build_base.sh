#!/bin/bash
# Builds the library of a given commit (default HEAD) as nind_denoise_b200/libnind_b200_base.so and the probe as
# tools/probe_base, for same-box A/B runs against the working tree:  NIND_LIB=nind_denoise_b200/libnind_b200_base.so python bench.py ...
set -e
REV=${1:-HEAD}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
T=/tmp/nind_base_$$
mkdir -p $T/nind_denoise_b200/csrc $T/include $T/tools $T/obj
for f in $(git -C $ROOT ls-tree --name-only $REV nind_denoise_b200/csrc/ include/ tools/probe.cu); do git -C $ROOT show $REV:$f > $T/$f; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for src in $T/nind_denoise_b200/csrc/*.cu; do nvcc $FLAGS -c -o $T/obj/$(basename ${src%.cu}).o $src & done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/nind_denoise_b200/libnind_b200_base.so $T/obj/*.o
nvcc $FLAGS -o $ROOT/tools/probe_base $T/tools/probe.cu $T/obj/igemm_inst_*.o 2>/dev/null
rm -rf $T
ls -la $ROOT/nind_denoise_b200/libnind_b200_base.so $ROOT/tools/probe_base
