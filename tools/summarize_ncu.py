"""Turns gpurun_out/launches.csv + gpurun_out/prof_bench_raw.csv (tools/run_ncu_r2.sh) into
profiles/<prefix>_ncu_summary.md, profiles/<prefix>_ncu_launches_cs248.csv and profiles/ncu_traffic.json (what
bench.py's roofline.traffic reads, tagged with the hash of the kernel sources).   usage: summarize_ncu.py r02"""
import csv
import json
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nind_denoise_b200 import _build  # noqa: E402

prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
lpath = os.path.join(ROOT, "gpurun_out", "launches.csv")
rpath = os.path.join(ROOT, "gpurun_out", "prof_bench_raw.csv")
LAYERS = ["convs1.0", "convs1.2", "convs2.0", "convs2.2", "convs3.0", "convs3.2", "convs4.0", "convs4.2", "bottom.0",
          "bottom.2", "up1", "tconvs1.0", "tconvs1.2", "up2", "tconvs2.0", "tconvs2.2", "up3", "tconvs3.0", "tconvs3.2",
          "up4", "tconvs4.0", "tconvs4.2+head"]
out = [f"# ncu evidence ({prefix}; B200; `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --images 2`: "
       f"cs 248, 4 forwards of 133 crops per image)\n",
       f"Kernel sources: `{_build.build_key()[:16]}` (nind_denoise_b200/_build.py:build_key).  Captured by "
       "tools/run_ncu_r2.sh after the same command exited 0 without ncu.\n",
       "## Launch list of the two timed steps (`--metrics gpu__time_duration.sum --clock-control none --profile-from-start off`)\n"]
rows = [r for r in csv.reader(open(lpath)) if len(r) > 5]
hdr, data = rows[0], rows[1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = defaultdict(lambda: [0, 0.0]), 0.0
for r in data:
    name = r[ik].split("(")[0].replace("void ", "").replace("nind::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] == "ns" else (v * 1e3 if r[iu] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
out.append(f"{len(data)} launches, {tot / 1e3:.2f} ms of kernel time for 2 steps (cold-cache, serialised: compare shares, "
           "not absolutes)\n")
out.append("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t:.0f} | {100 * t / tot:.1f}% |")
conv = sum(t for k, (n, t) in agg.items() if "igemm" in k)
out.append(f"\nigemm_kernel share of kernel time: {100 * conv / tot:.1f}% (compare bench.py's live "
           "`roofline.kernel_share_of_step`)\n")

rows = list(csv.reader(open(rpath)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "ms", "time"), ("sm__cycles_elapsed.avg.per_second", "GHz", "SM clk"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%", "tensor pipe active"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "%", "smem wavefronts (tensor), of peak"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "%", "LSU data-pipe wavefronts, of peak"),
        ("dram__bytes_read.sum", "MB", "DRAM read"), ("dram__bytes_write.sum", "MB", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%", "DRAM throughput"),
        ("launch__registers_per_thread", "", "regs"), ("launch__block_size", "", "threads"),
        ("launch__shared_mem_per_block_dynamic", "KB", "smem")]
scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "Ghz": 1.0, "Mhz": 1e-3}
out.append("## `--set full` capture of the 22 conv launches of the first timed forward (133 crops; `--profile-from-start off -k regex:igemm -c 22`)\n")
out.append("| layer | kernel | " + " | ".join(f"{lab} ({u})" if u else lab for _, u, lab in want) + " |")
out.append("|---|---|" + "---|" * len(want))
traffic = {}
for li, r in enumerate(data[:len(LAYERS)]):
    kname = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("nind::", "")
    cells = []
    vals = {}
    for m, u, lab in want:
        if m not in hdr:
            cells.append("-")
            continue
        i = hdr.index(m)
        try:
            v = float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        except ValueError:
            cells.append(r[i])
            continue
        vals[m] = v
        cells.append(f"{v:.3f}" if v < 100 else f"{v:.1f}")
    traffic[LAYERS[li]] = (vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)) * 1e6
    out.append(f"| {LAYERS[li]} | `{kname}` | " + " | ".join(cells) + " |")
out.append("""
DRAM bytes per launch against the algorithmic bytes (bf16, each activation read once + written once) are in
`bench.py --layers` (GB column); the heaviest launch's measured traffic is what `roofline.traffic` reports.
The `.ncu-rep` is not committed (too large); regenerate with `gpurun -- bash tools/run_ncu_r2.sh`, then
`python tools/summarize_ncu.py r02`.
""")
open(os.path.join(ROOT, "profiles", f"{prefix}_ncu_summary.md"), "w").write("\n".join(out))
shutil.copy(lpath, os.path.join(ROOT, "profiles", f"{prefix}_ncu_launches_cs248.csv"))
heavy = max(traffic, key=lambda k: traffic[k]) if traffic else None
json.dump({"kernel_key": _build.build_key()[:16], "source": f"profiles/{prefix}_ncu_summary.md",
           "entries": {"UtNet_cs248_b133": {"layer": heavy, "dram_bytes": traffic.get(heavy),
                                            "per_layer_dram_bytes": traffic,
                                            "note": f"{heavy} launch of one 133-crop forward, dram__bytes_read.sum + "
                                                    f"dram__bytes_write.sum (ncu --set full)"}}},
          open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("\n".join(out))
