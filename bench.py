#!/usr/bin/env python
"""Benchmark of the hot path: megapixels/s of tiled UtNet denoising of a 24 MP synthetic image.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cs 248] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): the crops of ONE image are sharded by contiguous
raster range, every rank stitches its band and the bands are gathered to rank 0 over NCCL
(BASELINE.json configs[2]) — total work is fixed, so "scaling" is "strong".

One "step" = one whole 6000x4000 image through gather -> UtNet (bf16 tensor cores, fp32 accumulate)
-> trim/seam/stitch.  `value` has the image resident in HBM; `e2e` goes through the host-buffer
C-ABI entry (pinned host image in, pinned host image out, copies inside the timed region).
`--impl reference` times the reference algorithm's CPU implementation (the oracle port: torch fp32
on all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG = 6000, 4000
MP = W_IMG * H_IMG / 1e6
OL = 6
NETWORK = "UtNet"


def set_workload(network):
    """BASELINE.json configs[3]: UNet on a 45 MP (8256x5504) image, cs 512 / ucs 384 (extra, not the headline)."""
    global W_IMG, H_IMG, MP, NETWORK
    NETWORK = network
    if network == "UNet":
        W_IMG, H_IMG = 8256, 5504
        MP = W_IMG * H_IMG / 1e6


# dram__bytes_read.sum + dram__bytes_write.sum of the heaviest igemm launch, from the committed
# `ncu --set full` capture of this command (profiles/r01_v12_ncu_summary.md)
NCU_TRAFFIC = {(248, 168): 4.05e9}
NCU_TRAFFIC_NOTE = ("tconvs4.0 launch (168 crops, 128->64 ch @250^2): 4.05 GB measured vs 4.07 GB algorithmic "
                    "(bf16 in + out once); all five captured launches are within 4 % of algorithmic")


def workload(cs):
    if NETWORK == "UNet":
        return dict(cs=cs, ucs=(cs * 3) // 4, ol=OL)
    return dict(cs=cs, ucs=cs - 24, ol=OL)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        # keep the samples taken inside the timed region (the sampler is started before the warm-up so that
        # nvidia-smi's start-up / NVML initialisation cannot stall the timed launches)
        rows = [r[1:] for r in self.rows if len(r) >= 7 and (t0 is None or t0 - 0.05 <= r[0] <= t1 + 0.15)]
        if not rows:
            rows = [r[1:] for r in self.rows[-3:] if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(cs, sample_crops, threads=None):
    """Reference algorithm on the host cores (oracle port: torch fp32 CPU, reference loop semantics):
    gather + forward + trim/seam/add of `sample_crops` crops of the 24 MP image; MP/s is extrapolated
    to the whole image by crops (all crops cost the same)."""
    import numpy as np

    from oracle import geometry as og
    from oracle import nets as on

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    wl = workload(cs)
    g = og.crop_grid(W_IMG, H_IMG, wl["cs"], wl["ucs"], wl["ol"])
    sd = on.init_state_dict(NETWORK, seed=0)
    fwd = on.utnet_forward if NETWORK == "UtNet" else on.unet_forward
    img = np.random.default_rng(1).random((3, H_IMG, W_IMG), dtype=np.float32)
    idx = list(range(0, g.size, max(1, g.size // sample_crops)))[:sample_crops]
    with torch.no_grad():
        fwd(sd, torch.from_numpy(og.gather_crop(img, g, 0)).unsqueeze(0))  # warm-up
        t0 = time.perf_counter()
        out = np.zeros((3, H_IMG, W_IMG), np.float32)
        for i in idx:
            e = og.crop_entry(g, i)
            y = fwd(sd, torch.from_numpy(og.gather_crop(img, g, i)).unsqueeze(0))[0].numpy()
            xlo, ylo, xhi, yhi = e["usefuldim"]
            ax, ay = e["usefulstart"]
            t = y[:, ylo:yhi, xlo:xhi] * og.seam_weights(g, i)
            out[:, ay:ay + t.shape[1], ax:ax + t.shape[2]] += t
        dt = time.perf_counter() - t0
    mp_done = MP * len(idx) / g.size
    return mp_done / dt, dt, len(idx), g.size, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cs = args.cs
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(cs, 2)
    total_t = 0.0
    n_s = 0
    for _ in range(args.steps):
        r, dt, n_s, n_all, threads = cpu_reference_rate(cs, args.ref_sample)
        rates.append(r)
        total_t += dt
    val = statistics.mean(rates)
    wl = workload(cs)
    line = {
        "impl": "reference", "metric": "megapixels/sec denoised", "value": val, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{NETWORK}(funit 64) {W_IMG}x{H_IMG} synthetic image, cs {wl['cs']} ucs {wl['ucs']} overlap {OL}"},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": "port",
                         "sample": f"{n_s} of {n_all} crops per step (gather + fp32 forward + trim/seam/add), "
                                   f"extrapolated by crop count"},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cs", type=int, default=0, help="crop size (legal UtNet sizes: 120, 248 [default], 504, 1016; "
                                                       "UNet: multiple of 16, default 512)")
    ap.add_argument("--network", default="UtNet", choices=["UtNet", "UNet"])
    ap.add_argument("--batch", type=int, default=0, help="crops per forward (0 = auto)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=0, help="crops per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="rows", choices=["rows", "bands"],
                    help="N>1: seam exchange + owned-row gather (default) or whole-band gather summed on rank 0")
    ap.add_argument("--layers", action="store_true", help="print the per-layer timing table to stderr")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="nind_set_option knob for A/B runs (e.g. pair64=0); not used for the headline")
    args = ap.parse_args()
    set_workload(args.network)
    if args.cs <= 0:
        args.cs = 248 if args.network == "UtNet" else 512
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.ref_sample <= 0:
        args.ref_sample = 48 if args.cs <= 264 else 6
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    import nind_denoise_b200 as nb
    from nind_denoise_b200 import _capi
    from nind_denoise_b200.tiler import _band, default_batch
    from nind_denoise_b200.flops import unet_flops, utnet_flops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload(args.cs)
    cs, ucs, ol = wl["cs"], wl["ucs"], wl["ol"]

    torch.manual_seed(0)  # default init == the reference class's default init under the same seed
    model = (nb.UtNet() if NETWORK == "UtNet" else nb.UNet()).to(dev).eval()
    for kv in args.opt:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    g = torch.Generator(device="cpu").manual_seed(1)
    img_host = torch.rand((3, H_IMG, W_IMG), generator=g).pin_memory()
    out_host = torch.empty_like(img_host).pin_memory()
    img = img_host.to(dev)
    n = nb.n_crops(W_IMG, H_IMG, cs, ucs, ol)
    ranges = nb.shard_ranges(n, world)
    cb, ce = ranges[rank]
    batch = args.batch or default_batch(ce - cb, cs, -(-(W_IMG - ucs) // (ucs - ol)) + 1)
    lib = _capi.lib()

    def step():
        if world == 1:
            return nb.denoise_tiled(img, model, cs, ucs, ol, batch=batch)
        return nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch, mode=args.gather)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    sync_all()
    l0 = lib.nind_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_start = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    sync_all()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.nind_kernel_launches() - l0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None

    # ---- end to end through the host-buffer entry point (N = 1) / host image in + out (N > 1)
    # N > 1: one shared, page-locked host image that every rank writes the rows it owns into (N PCIe links)
    out_shared = nb.SharedHostImage((3, H_IMG, W_IMG)) if world > 1 else None

    def e2e_step():
        if world == 1:
            nb.denoise_tiled_host(img_host, model, cs, ucs, ol, batch=batch, out=out_host)
        else:
            nb.denoise_tiled_distributed_host(img_host, model, cs, ucs, ol, batch=batch, out=out_shared)

    e2e_step()
    e2e_step()  # the host pipeline alternates between two device slots: warm both
    sync_all()
    e2e_steps = max(2, min(args.steps, 10))
    t0 = time.perf_counter()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms = e0.elapsed_time(e1)
    t2 = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item())
    if world == 1:
        h2d = img_host.numel() * 4
    else:  # every rank uploads only the rows its crops read
        from nind_denoise_b200.tiler import rows_needed
        h2d = sum((lambda r: (r[1] - r[0]) * W_IMG * 12)(rows_needed(W_IMG, H_IMG, cs, ucs, ol, a, b))
                  for a, b in ranges if b > a)
    d2h = out_host.numel() * 4

    # ---- throughput mode (BASELINE configs[4] in miniature): a stream of images through the async host entry
    # N > 1: every rank streams whole images on its own GPU (replicas, no collective); aggregate over ranks.
    n_img = 6
    full_batch = args.batch or default_batch(n, cs, -(-(W_IMG - ucs) // (ucs - ol)) + 1)
    outs2 = [out_host, torch.empty_like(img_host).pin_memory()]
    nb.denoise_images_host([img_host] * 2, model, cs, ucs, ol, batch=full_batch, outs=outs2)
    sync_all()
    t0 = time.perf_counter()
    nb.denoise_images_host([img_host] * n_img, model, cs, ucs, ol, batch=full_batch,
                           outs=[outs2[i & 1] for i in range(n_img)])
    torch.cuda.synchronize()
    t3 = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    thr = {"value": MP * n_img * world / float(t3.item()), "unit": "MP/s", "images": n_img * world,
           "note": "stream of images, host buffers in/out, H2D/D2H of neighbouring images overlapped"
                   + (f"; {n_img} whole images per GPU, {world} independent replicas" if world > 1 else "")}

    # ---- roofline of the dominant kernel (igemm conv): per-layer CUDA-event times over one image
    flops_image = (utnet_flops(cs) if NETWORK == "UtNet" else unet_flops(cs)) * n
    pk = peaks()
    roof = None
    layer_rows = []
    if rank == 0:
        _capi.check(lib.nind_set_timing(model.native_handle(), 1))
        import ctypes as C
        agg = {}
        nb_ = 0
        for i0 in range(cb, ce, batch):
            i1 = min(ce, i0 + batch)
            _band(model, img, cs, ucs, ol, i0, i1, batch)
            cnt = C.c_int()
            names = (C.c_char_p * 128)()
            tms = (C.c_float * 128)()
            fl = (C.c_double * 128)()
            _capi.check(lib.nind_get_layer_times(model.native_handle(), 128, names, tms, fl, C.byref(cnt)))
            by = (C.c_double * 128)()
            _capi.check(lib.nind_get_layer_bytes(model.native_handle(), 128, by, C.byref(cnt)))
            for k in range(cnt.value):
                a = agg.setdefault(names[k].decode(), [0.0, 0.0, 0, 0.0])
                a[0] += tms[k]; a[1] += fl[k]; a[2] += 1; a[3] += by[k]
            nb_ += 1
        _capi.check(lib.nind_set_timing(model.native_handle(), 0))
        conv_ms = sum(v[0] for k, v in agg.items() if v[1] > 0)
        conv_fl = sum(v[1] for k, v in agg.items() if v[1] > 0)
        all_ms = sum(v[0] for v in agg.values())
        n_launch = sum(v[2] for k, v in agg.items() if v[1] > 0)
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": "nind::igemm_kernel<N_TILE> (all conv layers)",
                "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                "frac_of_burst": achieved / pk["burst"], "peak_source": pk["source"] + " (bf16_tflops_sustained)",
                "traffic": NCU_TRAFFIC.get((cs, batch)), "traffic_note": NCU_TRAFFIC_NOTE if (cs, batch) in NCU_TRAFFIC else None,
                "launches": n_launch, "avg_launch_ms": conv_ms / max(1, n_launch),
                "kernel_share_of_step": conv_ms / all_ms if all_ms else None}
        layer_rows = sorted(((k, v[0], v[1], v[3]) for k, v in agg.items()), key=lambda r: -r[1])
        # HBM-bound kernels of the step (gather, 2x2/s2 upsamplers, first layer): achieved GB/s
        if True:
            roof["memory_bound_kernels"] = {
                k: {"GB/s": v[3] / (v[0] * 1e-3) / 1e9, "frac_of_hbm_peak": v[3] / (v[0] * 1e-3) / 1e9 / pk["hbm"]}
                for k, v in agg.items() if k in ("gather+pad8", "up4", "up3", "convs1.0", "inc.conv.conv.0", "up4.up", "up3.up") and v[0] > 0}
        if args.layers:
            for k, tm, f, b_ in layer_rows:
                print(f"{k:28s} {tm:8.3f} ms  {f / 1e9:10.1f} GFLOP  {f / (tm * 1e-3) / 1e12 if tm else 0:8.1f} TF/s"
                      f"  {b_ / 1e9:8.2f} GB  {b_ / (tm * 1e-3) / 1e9 if tm else 0:8.0f} GB/s", file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, dt, n_s, n_all, threads = cpu_reference_rate(cs, 4 * args.ref_sample)  # ~10-20 s of CPU work
        cpu = {"value": r, "unit": "MP/s", "cores": threads, "kind": "port",
               "sample": f"{n_s} of {n_all} crops (gather + fp32 torch forward + trim/seam/add, {dt:.1f} s), "
                         f"extrapolated by crop count"}

    if rank == 0:
        value = MP * args.steps / (ms * 1e-3)
        line = {
            "metric": "megapixels/sec denoised", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{NETWORK}(funit 64, random init) {W_IMG}x{H_IMG} synthetic image, cs {cs} ucs {ucs} overlap {ol} "
                                   f"-> {n} crops, batch {batch} crops/forward; crops sharded over {world} GPU(s)"
                                   + (f", NCCL seam exchange between neighbours + send/recv gather of owned rows to rank 0 ({args.gather})"
                                      if world > 1 else ""),
                       "l2": "inputs larger than L2 (288 MB image, >1 GB activation arena per batch)",
                       "algorithmic_tflop_per_step": flops_image / 1e12},
            "pct_of_bf16_peak": {"sustained": flops_image / (ms / args.steps * 1e-3) / 1e12 / (pk["sustained"] * world),
                                 "burst": flops_image / (ms / args.steps * 1e-3) / 1e12 / (pk["burst"] * world)},
            "e2e": {"value": MP * e2e_steps / (e2e_ms * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "throughput_mode": thr,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
