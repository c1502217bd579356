#!/bin/bash
# one-GPU regression: tests, smoke, default bench (cs 248) with CPU baseline, cs 504, reference arm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --layers > gpurun_out/bench_248.json 2> gpurun_out/bench_248.err; cut -c1-300 gpurun_out/bench_248.json
python bench.py --cs 504 --layers --no-cpu-baseline > gpurun_out/bench_504.json 2> gpurun_out/bench_504.err; cut -c1-300 gpurun_out/bench_504.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-700
