"""Turns gpurun_out/launches.csv + gpurun_out/prof_bench.ncu-rep (tools/run_ncu_bench.sh) into
profiles/<prefix>_ncu_summary.md and copies the launch list next to it.   usage: summarize_ncu.py r01"""
import csv, io, os, shutil, subprocess, sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r01"
lpath = os.path.join(ROOT, "gpurun_out", "launches.csv")
rpath = os.path.join(ROOT, "gpurun_out", "prof_bench.ncu-rep")
out = []
out.append(f"# ncu evidence ({prefix}; B200; `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`: cs 248, batch 168)\n")
out.append("Captured by tools/run_ncu_launches.sh and tools/run_ncu_full.sh after the same command exited 0 without ncu.\n")
out.append("## Launch list of the two timed steps (`--metrics gpu__time_duration.sum --clock-control none -s 279 -c 186`)\n")
rows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = defaultdict(lambda: [0, 0.0]), 0.0
for r in data:
    name = r[ik].split("(")[0].replace("void ", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v; tot += v
out.append(f"{len(data)} launches, {tot / 1e3:.2f} ms of kernel time for 2 steps (cold-cache, serialised: compare shares, not absolutes)\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% |")
conv = sum(t for k, (n, t) in agg.items() if "igemm" in k)
out.append(f"\nigemm_kernel share of kernel time: {100 * conv / tot:.1f}% (compare bench.py's live `roofline.kernel_share_of_step`)\n")
raw = subprocess.run(["ncu", "-i", rpath, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
names = ["tconvs3.0 (256->128, 3x3)", "tconvs3.2 (128->128, 3x3)", "up4 (128->4x64, 2x2 s2)", "tconvs4.0 (128->64, 3x3)",
         "tconvs4.2+head (64->64, 3x3, +1x1)"]
want = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts by tensor core, of peak"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
        ("smsp__inst_executed.sum", "warp instructions executed"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__block_size", "threads/CTA"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block")]
out.append("## `--set full` capture of five conv launches of the first timed batch (`-k regex:igemm -s 281 -c 5`)\n")
out.append("| metric | " + " | ".join(names[: len(data)]) + " |\n|---|" + "---|" * len(data))
out.append("| kernel | " + " | ".join("`" + r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") + "`" for r in data) + " |")
for m, label in want:
    if m in hdr:
        i = hdr.index(m)
        out.append(f"| {label} ({units[i]}) | " + " | ".join(r[i] for r in data) + " |")
out.append("""
Algorithmic bytes per launch (168 crops of 248, bf16, each activation read once + written once):
tconvs3.0 1.32 + 0.64 GB, tconvs3.2 0.68 + 0.66 GB, up4 0.66 + 1.32 GB, tconvs4.0 2.73 + 1.34 GB, tconvs4.2+head
1.39 GB + 0.12 GB (fp32 planar image) — compare the
DRAM read / write rows: no re-reads (the 9 taps are served from the shared-memory patch, weights from
shared memory / L2).  The `.ncu-rep` (31 MB) is not committed; regenerate with
`gpurun -- bash tools/run_ncu_launches.sh`, then `tools/run_ncu_full.sh` and `python tools/summarize_ncu.py <prefix>`.
""")
open(os.path.join(ROOT, "profiles", f"{prefix}_ncu_summary.md"), "w").write("\n".join(out))
shutil.copy(lpath, os.path.join(ROOT, "profiles", f"{prefix}_ncu_launches_cs248.csv"))
print("\n".join(out))
