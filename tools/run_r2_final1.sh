#!/bin/bash
# round 2, final single-GPU regression: GPU tests, smoke, default bench (with cpu_baseline), reference arm, cs 504 and
# UNet lines, then the ncu evidence of the bench command (tools/run_ncu_r2.sh).
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default)"
python bench.py --layers 2> gpurun_out/r2z_layers.txt > gpurun_out/r2z_bench.json; cut -c1-250 gpurun_out/r2z_bench.json
echo "=== bench --impl reference"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2z_reference.json
echo "=== bench --cs 504 / --network UNet"
python bench.py --steps 5 --no-cpu-baseline --cs 504 > gpurun_out/r2z_bench_504.json; cut -c1-200 gpurun_out/r2z_bench_504.json
python bench.py --steps 3 --no-cpu-baseline --network UNet > gpurun_out/r2z_bench_unet.json; cut -c1-200 gpurun_out/r2z_bench_unet.json
echo "=== ncu"
bash tools/run_ncu_r2.sh
