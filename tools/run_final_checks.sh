#!/bin/bash
# one-GPU regression: tests, smoke, default bench (cs 248) with CPU baseline, cs 504, UNet, reference arm
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --layers > gpurun_out/bench_248.json 2> gpurun_out/bench_248.err; cat gpurun_out/bench_248.json | cut -c1-3000
python bench.py --cs 504 --layers --no-cpu-baseline > gpurun_out/bench_504.json 2> gpurun_out/bench_504.err; cat gpurun_out/bench_504.json | cut -c1-700
python bench.py --network UNet --steps 5 --layers --no-cpu-baseline > gpurun_out/bench_unet.json 2> gpurun_out/bench_unet.err; cat gpurun_out/bench_unet.json | cut -c1-700
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-700
