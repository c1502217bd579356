#!/usr/bin/env python
"""Benchmark of the hot path: megapixels/s of tiled UtNet denoising of a 24 MP synthetic image.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cs 248] [--network UtNet|UNet] [--impl reference]
                    [--images M] [--sweep]

N > 1 is launched by torchrun (one rank per GPU): the crops of ONE image are sharded by contiguous
raster range, every rank stitches its band and the bands are gathered to rank 0 over NCCL
(BASELINE.json configs[2]) — total work is fixed, so "scaling" is "strong".

One "step" = one whole 6000x4000 image through gather -> UtNet (bf16 tensor cores, fp32 accumulate)
-> trim/seam/stitch.  `value` has the image resident in HBM; `e2e` goes through the host-buffer
C-ABI entry (pinned host image in, pinned host image out, copies inside the timed region).
`--network UNet` is BASELINE configs[3] (45 MP image, cs 512); `--images M` / `--sweep` are configs[4]
(a stream of M images per GPU through the asynchronous host entry; crop-size x overlap sweep).
`--impl reference` times the reference's own CPU implementation (oracle/_ref: the reference's classes, taken
by oracle/make_ref.py; else the oracle port) on a bounded sample of the same workload.

Every run also prints a `parity` block, computed outside the timed regions: the stitched image against the
fp32 CPU oracle on sampled crops, and at N > 1 the sharded results against a single-GPU run of the same image.
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


def workload(cs, ol=OL):
    if NETWORK == "UNet":
        return dict(cs=cs, ucs=(cs * 3) // 4, ol=ol)
    return dict(cs=cs, ucs=cs - 24, ol=ol)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


def ncu_traffic(cs, batch):
    """dram__bytes_read + dram__bytes_write of the heaviest conv launch, from the committed `ncu --set full`
    summary of this command (profiles/ncu_traffic.json, written by tools/summarize_ncu.py together with the
    hash of the kernel sources it was captured from).  A capture of other sources is reported as stale."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None, "no ncu capture committed"
    d = json.load(open(path))
    from nind_denoise_b200 import _build
    cur = _build.build_key()[:16]
    e = d.get("entries", {}).get(f"{NETWORK}_cs{cs}_b{batch}")
    if e is None:
        return None, f"no capture for {NETWORK} cs {cs} batch {batch}"
    note = e.get("note", "")
    if d.get("kernel_key") != cur:
        note += f" [captured from kernel sources {d.get('kernel_key')}, current {cur}: stale]"
    return e.get("dram_bytes"), note


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled during the timed region (B200_PROFILING.md's clocks line).

    Read through NVML in this process — the library `nvidia-smi` itself reads, two calls per sample every 100 ms —
    because a looping `nvidia-smi --query-gpu=... -lms 100` next to the bench occasionally stalls the GPU for tens of
    milliseconds (one 60 ms step among sixty 36 ms ones with it, none without: per-step stamps under
    NIND_BENCH_DIAG=1).  Falls back to the `nvidia-smi` loop when pynvml cannot be imported."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.nvml, self.stop_flag, self.source = [], None, index, None, False, None

    def _nvml_loop(self):
        n = self.nvml
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append([time.time(), str(sm), str(self.sm_max)] +
                                 ["Active" if r & b else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.source = pynvml, "nvml"
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self, t0=None, t1=None):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        time.sleep(0.15)
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        # keep the samples taken inside the timed region (the sampler is started before the warm-up so that
        # its start-up / NVML initialisation cannot stall the timed launches)
        rows = [r[1:] for r in self.rows if len(r) >= 7 and (t0 is None or t0 - 0.05 <= r[0] <= t1 + 0.15)]
        if not rows:
            rows = [r[1:] for r in self.rows[-3:] if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for r in rows for i in range(4) if r[2 + i].startswith("Active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


def sample_crops(nx, ny, k):
    """Crop indices for the parity check: the four corners, edge mid-points and interior crops."""
    cand = [0, nx * ny - 1, (ny // 2) * nx + nx // 2, nx - 1, (ny - 1) * nx, nx // 2, (ny // 2) * nx,
            (ny // 2) * nx + nx - 1, (ny - 1) * nx + nx // 2, (ny // 3) * nx + (2 * nx) // 3]
    out = []
    for c in cand:
        if 0 <= c < nx * ny and c not in out:
            out.append(c)
    return out[:k]


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(cs, n_sample, threads=None):
    """The reference's implementation of the path on the host cores: crop gather (OneImageDS.__getitem__), fp32
    forward, trim / seam / overlap-add, for a sample of crops; rate extrapolated by crop count.  Runs the
    reference's own classes when oracle/_ref is present (kind "reference"), else the oracle port (kind "port").
    Runs under torch.no_grad() — the stock script builds (and drops) an autograd graph per crop — which favours
    the reference."""
    import numpy as np

    from oracle import geometry as og
    from oracle import make_ref
    from oracle import nets as on

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    wl = workload(cs)
    g = og.crop_grid(W_IMG, H_IMG, wl["cs"], wl["ucs"], wl["ol"])
    img = np.random.default_rng(1).random((3, H_IMG, W_IMG), dtype=np.float32)
    idx = list(range(0, g.size, max(1, g.size // n_sample)))[:n_sample]
    ol, ucs = wl["ol"], wl["ucs"]
    if make_ref.available():
        kind = "reference"
        UtNet, UNet, dataset = make_ref.load()
        torch.manual_seed(0)
        model = (UtNet() if NETWORK == "UtNet" else UNet()).eval()
        ds = dataset(img, wl["cs"], ucs, ol)
        with torch.no_grad():
            model(ds[0][0].unsqueeze(0))  # warm-up
            t0 = time.perf_counter()
            newimg = torch.zeros(3, H_IMG, W_IMG, dtype=torch.float32)
            for i in idx:  # denoise_image.py:240-267, batch_size 1
                y, ud, us = ds[i]
                x = model(y.unsqueeze(0))
                t = x[0][:, ud[1]:ud[3], ud[0]:ud[2]].cpu().detach()
                ax, ay = tuple(us.tolist())
                if ax != 0:
                    t[:, :, 0:ol] = t[:, :, 0:ol].div(2)
                if ay != 0:
                    t[:, 0:ol, :] = t[:, 0:ol, :].div(2)
                if ax + ucs < W_IMG and ol:
                    t[:, :, -ol:] = t[:, :, -ol:].div(2)
                if ay + ucs < H_IMG and ol:
                    t[:, -ol:, :] = t[:, -ol:, :].div(2)
                newimg[:, ay:ay + t.shape[1], ax:ax + t.shape[2]] += t
            dt = time.perf_counter() - t0
    else:
        kind = "port"
        sd = on.init_state_dict(NETWORK, seed=0)
        fwd = on.utnet_forward if NETWORK == "UtNet" else on.unet_forward
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
    return mp_done / dt, dt, len(idx), g.size, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cs = args.cs
    rates = []
    for _ in range(args.warmup):
        cpu_reference_rate(cs, 2)
    total_t, n_s, n_all, threads, kind = 0.0, 0, 0, os.cpu_count(), "port"
    for _ in range(args.steps):
        r, dt, n_s, n_all, threads, kind = cpu_reference_rate(cs, args.ref_sample)
        rates.append(r)
        total_t += dt
    val = statistics.mean(rates)
    wl = workload(cs)
    line = {
        "impl": "reference", "metric": "megapixels/sec denoised", "value": val, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{NETWORK}(funit 64, random init) {W_IMG}x{H_IMG} synthetic image, cs {wl['cs']} "
                               f"ucs {wl['ucs']} overlap {wl['ol']}"},
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": kind,
                         "sample": f"{n_s} of {n_all} crops per step (OneImageDS gather + fp32 forward + trim/seam/add, "
                                   f"torch.no_grad), extrapolated by crop count"},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------- parity block
def oracle_parity(img_host, outs, cs, ucs, ol, k):
    """Stitched GPU image(s) against the fp32 CPU oracle on `k` sampled crops: the pixels a crop owns exclusively
    (its useful area minus the seam bands shared with neighbours) must equal the oracle's forward of that crop.
    Returns max abs error, error / sigma_out and PSNR(got, oracle) per output in `outs` (dict name -> CPU tensor)."""
    import math

    import numpy as np

    from oracle import geometry as og
    from oracle import nets as on

    g = og.crop_grid(W_IMG, H_IMG, cs, ucs, ol)
    sd = on.init_state_dict(NETWORK, seed=0)
    fwd = on.utnet_forward if NETWORK == "UtNet" else on.unet_forward
    img = img_host.numpy()
    idx = sample_crops(g.nx, g.ny, k)
    torch.set_num_threads(os.cpu_count())
    stats = {name: dict(err=0.0, se=0.0, n=0) for name in outs}
    ref_all = []
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in idx:
            e = og.crop_entry(g, i)
            y = fwd(sd, torch.from_numpy(og.gather_crop(img, g, i)).unsqueeze(0))[0].numpy()
            xlo, ylo, xhi, yhi = e["usefuldim"]
            ax, ay = e["usefulstart"]
            h, w = yhi - ylo, xhi - xlo
            l = ol if ax != 0 else 0
            t = ol if ay != 0 else 0
            r = ol if (ax + ucs < W_IMG and ol) else 0
            b = ol if (ay + ucs < H_IMG and ol) else 0
            ref = y[:, ylo + t:yhi - b, xlo + l:xhi - r]
            ref_all.append(ref.ravel())
            for name, o in outs.items():
                got = o[:, ay + t:ay + h - b, ax + l:ax + w - r].numpy()
                d = np.abs(got.astype(np.float64) - ref)
                st = stats[name]
                st["err"] = max(st["err"], float(d.max()))
                st["se"] += float((d ** 2).sum())
                st["n"] += d.size
    sigma = float(np.concatenate(ref_all).std())
    res = {"crops": idx, "oracle_seconds": round(time.perf_counter() - t0, 2), "sigma_out": sigma,
           "tolerance": {"max_abs": 2e-2, "max_abs_over_sigma": 0.25}}
    ok = True
    for name, st in stats.items():
        mse = st["se"] / max(1, st["n"])
        res[name] = {"max_abs": st["err"], "max_abs_over_sigma": st["err"] / sigma,
                     "psnr_vs_oracle_db": 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)}
        ok = ok and st["err"] <= 2e-2 and st["err"] <= 0.25 * sigma
    res["pass"] = bool(ok)
    return res


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
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle / distributed parity block")
    ap.add_argument("--parity-crops", type=int, default=0, help="crops checked against the oracle (0 = auto)")
    ap.add_argument("--gather", default="peer", choices=["peer", "rows", "bands"],
                    help="N>1: peer-memory DMA gather overlapped with compute (default), NCCL seam exchange + "
                         "owned-row send/recv, or whole-band gather summed on rank 0")
    ap.add_argument("--images", type=int, default=6, help="throughput mode: images streamed per GPU")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[4]: throughput mode over cs {120,248,504,1016} x overlap {0,6,16,32}")
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
    from nind_denoise_b200.flops import unet_flops, utnet_flops
    from nind_denoise_b200.tiler import _band, default_batch, rows_needed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one process per GPU: run on (and allocate the pinned host buffers from) the GPU's own NUMA node
        if os.environ.get("NIND_NO_NUMA_BIND") is None:
            nb.bind_host_to_gpu(local)
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
    out_host2 = torch.empty_like(img_host).pin_memory()
    img = img_host.to(dev)
    n = nb.n_crops(W_IMG, H_IMG, cs, ucs, ol)
    nx = -(-(W_IMG - ucs) // (ucs - ol)) + 1
    ranges = nb.shard_ranges(n, world)
    cb, ce = ranges[rank]
    batch = args.batch or default_batch(ce - cb, cs, nx)
    lib = _capi.lib()
    pk = peaks()
    flops_crop = utnet_flops(cs) if NETWORK == "UtNet" else unet_flops(cs)
    flops_image = flops_crop * n

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def stream_images(n_img, cs_, ucs_, ol_, batch_):
        """Throughput mode: n_img images per GPU through the asynchronous host entry; seconds (max over ranks)."""
        outs2 = [out_host, out_host2]
        nb.denoise_images_host([img_host] * 2, model, cs_, ucs_, ol_, batch=batch_, outs=outs2)
        sync_all()
        t0 = time.perf_counter()
        nb.denoise_images_host([img_host] * n_img, model, cs_, ucs_, ol_, batch=batch_,
                               outs=[outs2[i & 1] for i in range(n_img)])
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    # ------------------------------------------------------------------ --sweep: configs[4] only
    if args.sweep:
        rows = []
        for cs_s in (120, 248, 504, 1016):
            for ol_s in (0, 6, 16, 32):
                w = workload(cs_s, ol_s)
                n_s = nb.n_crops(W_IMG, H_IMG, w["cs"], w["ucs"], w["ol"])
                nx_s = -(-(W_IMG - w["ucs"]) // (w["ucs"] - w["ol"])) + 1
                b_s = default_batch(n_s, w["cs"], nx_s)
                sec = stream_images(args.images, w["cs"], w["ucs"], w["ol"], b_s)
                fl = (utnet_flops(w["cs"]) if NETWORK == "UtNet" else unet_flops(w["cs"])) * n_s
                rows.append({"cs": w["cs"], "ucs": w["ucs"], "overlap": w["ol"], "crops": n_s, "batch": b_s,
                             "MP/s": MP * args.images * world / sec,
                             "frac_of_sustained_bf16": fl * args.images / sec / 1e12 / pk["sustained"]})
                if rank == 0:
                    print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
        if rank == 0:
            head = next(r for r in rows if r["cs"] == 248 and r["overlap"] == 6)
            print(json.dumps({
                "metric": "megapixels/sec denoised", "value": head["MP/s"], "unit": "MP/s", "n_gpus": world,
                "steps": args.images, "warmup": 2, "ms_per_step": 1e3 * MP * world / head["MP/s"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"throughput mode: {args.images} x {W_IMG}x{H_IMG} images per GPU on {world} "
                                       f"replica(s), host buffers in/out; crop-size x overlap sweep (value = cs 248, overlap 6)"},
                "sweep": rows}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    def step():
        if world == 1:
            return nb.denoise_tiled(img, model, cs, ucs, ol, batch=batch)
        return nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch, mode=args.gather)

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
    torch.cuda.profiler.start()  # no-op unless run under `ncu --profile-from-start off` (tools/run_ncu_r2.sh)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        out = step()
        marks[i].record()   # per-step stamps (no synchronisation): a stall shows up as one long step
    e1.record()
    sync_all()
    torch.cuda.profiler.stop()
    t_end = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    step_ms = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    launches = lib.nind_kernel_launches() - l0
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    if os.environ.get("NIND_BENCH_DIAG") and rank == 0:
        print("diag: per-step ms with the clock sampler running: " + " ".join(f"{v:.2f}" for v in step_ms), file=sys.stderr)
        m2 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        m2[0].record()
        for i in range(args.steps):
            step()
            m2[i + 1].record()
        torch.cuda.synchronize()
        print("diag: per-step ms after the sampler was stopped:      " +
              " ".join(f"{a.elapsed_time(b):.2f}" for a, b in zip(m2, m2[1:])), file=sys.stderr)

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
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    if world == 1:
        h2d = img_host.numel() * 4
    else:  # every rank uploads only the rows its crops read
        h2d = sum((lambda r: (r[1] - r[0]) * W_IMG * 12)(rows_needed(W_IMG, H_IMG, cs, ucs, ol, a, b))
                  for a, b in ranges if b > a)
    d2h = out_host.numel() * 4

    # ---- parity (outside every timed region)
    parity = None
    if not args.no_parity:
        parity = {}
        if world > 1 and rank == 0:
            # sharded == single GPU: rank 0 runs the whole image alone and compares both sharded results with it
            single = nb.denoise_tiled(img, model, cs, ucs, ol, batch=default_batch(n, cs, nx))
            dvs = {"device_path": args.gather, "device_path_max_abs": float((out - single).abs().max()),
                   "shared_host_path_max_abs": float((out_shared.tensor - single.cpu()).abs().max()),
                   "tolerance": 1e-6,
                   "note": "4-way seam corners may associate (a+b)+(c+d) instead of ((a+b)+c)+d across rank boundaries"}
            dvs["pass"] = bool(max(dvs["device_path_max_abs"], dvs["shared_host_path_max_abs"]) <= 1e-6)
            parity["distributed_vs_single"] = dvs
            del single
        if rank == 0:
            k = args.parity_crops or (6 if cs <= 264 else 4)
            outs = {"device_resident": out.cpu(),
                    "host_e2e": (out_host if world == 1 else out_shared.tensor).clone()}
            parity["vs_oracle"] = oracle_parity(img_host, outs, cs, ucs, ol, k)
            del outs
    sync_all()

    # ---- throughput mode (BASELINE configs[4]): a stream of images through the async host entry
    # N > 1: every rank streams whole images on its own GPU (replicas, no collective); aggregate over ranks.
    n_img = max(2, args.images)
    full_batch = args.batch or default_batch(n, cs, nx)
    sec = stream_images(n_img, cs, ucs, ol, full_batch)
    thr = {"value": MP * n_img * world / sec, "unit": "MP/s", "images": n_img * world,
           "note": "stream of images, host buffers in/out, H2D/D2H of neighbouring images overlapped"
                   + (f"; {n_img} whole images per GPU, {world} independent replicas" if world > 1 else "")}

    # ---- roofline of the dominant kernel (igemm conv): per-layer CUDA-event times over one image
    roof = None
    if rank == 0:
        import ctypes as C

        from nind_denoise_b200 import _build

        _capi.check(lib.nind_set_timing(model.native_handle(), 1))
        agg = {}
        # the same forwards the timed step ran: balanced batches of the rank's crop range
        k_steps = -(-(ce - cb) // batch)
        for i in range(k_steps):
            i0, i1 = cb + (ce - cb) * i // k_steps, cb + (ce - cb) * (i + 1) // k_steps
            _band(model, img, cs, ucs, ol, i0, i1, i1 - i0)
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
        _capi.check(lib.nind_set_timing(model.native_handle(), 0))
        conv_ms = sum(v[0] for v in agg.values() if v[1] > 0)
        conv_fl_exec = sum(v[1] for v in agg.values() if v[1] > 0)
        all_ms = sum(v[0] for v in agg.values())
        n_launch = sum(v[2] for v in agg.values() if v[1] > 0)
        # SURVEY §8d convention: algorithmic FLOPs (Conv2d output pixels, ConvTranspose2d input pixels) of the
        # crops this rank ran, over the CUDA-event time of its conv launches
        conv_fl_alg = flops_crop * (ce - cb)
        achieved = conv_fl_alg / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        traffic, traffic_note = ncu_traffic(cs, -(-(ce - cb) // k_steps))  # the forwards are balanced: 4 x 133 crops
        roof = {"bound": "tensor", "kernel": "nind::igemm_kernel<N_TILE, TPS, CG, C8, PM> (all conv layers)",
                "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": achieved / pk["sustained"],
                "frac_of_burst": achieved / pk["burst"], "peak_source": pk["source"] + " (bf16_tflops_sustained)",
                "flops": "algorithmic (SURVEY 8d: Conv2d output pixels, ConvTranspose2d input pixels)",
                "algorithmic_tflop": conv_fl_alg / 1e12, "executed_tflop": conv_fl_exec / 1e12,
                "achieved_executed": conv_fl_exec / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0,
                "traffic": traffic, "traffic_note": traffic_note,
                "launches": n_launch, "avg_launch_ms": conv_ms / max(1, n_launch), "kernel_ms_per_step": conv_ms,
                "kernel_share_of_step": conv_ms / all_ms if all_ms else None,
                "kernel_sources": _build.build_key()[:16]}
        layer_rows = sorted(((k, v[0], v[1], v[3]) for k, v in agg.items()), key=lambda r: -r[1])
        roof["layers"] = {k: {"ms": round(tm, 4), "executed_TFLOP/s": round(f / (tm * 1e-3) / 1e12, 1) if tm else 0.0}
                          for k, tm, f, _ in layer_rows[:8]}
        # HBM-bound kernels of the step (gather, 2x2/s2 upsamplers, first layer): achieved GB/s
        mem_names = ("gather+pad8", "up4", "up3", "convs1.0", "inc.conv.conv.0", "up4.up", "up3.up")
        roof["memory_bound_kernels"] = {
            k: {"GB/s": v[3] / (v[0] * 1e-3) / 1e9, "frac_of_hbm_peak": v[3] / (v[0] * 1e-3) / 1e9 / pk["hbm"]}
            for k, v in agg.items() if k in mem_names and v[0] > 0}
        roof["ms_outside_conv_kernels"] = ms / args.steps - conv_ms
        if args.layers:
            for k, tm, f, b_ in layer_rows:
                print(f"{k:28s} {tm:8.3f} ms  {f / 1e9:10.1f} GFLOP  {f / (tm * 1e-3) / 1e12 if tm else 0:8.1f} TF/s"
                      f"  {b_ / 1e9:8.2f} GB  {b_ / (tm * 1e-3) / 1e9 if tm else 0:8.0f} GB/s", file=sys.stderr)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r, dt, n_s, n_all, threads, kind = cpu_reference_rate(cs, 4 * args.ref_sample)  # ~10-30 s of CPU work
        cpu = {"value": r, "unit": "MP/s", "cores": threads, "kind": kind,
               "sample": f"{n_s} of {n_all} crops (OneImageDS gather + fp32 torch forward + trim/seam/add, {dt:.1f} s, "
                         f"torch.no_grad), extrapolated by crop count"}

    if rank == 0:
        value = MP * args.steps / (ms * 1e-3)
        line = {
            "metric": "megapixels/sec denoised", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{NETWORK}(funit 64, random init) {W_IMG}x{H_IMG} synthetic image, cs {cs} ucs {ucs} overlap {ol} "
                                   f"-> {n} crops, batch {batch} crops/forward; crops sharded over {world} GPU(s)"
                                   + ((", owned rows copied to rank 0 by peer DMA over NVLink as they become final, NCCL for the "
                                       "closing synchronisation (peer)" if args.gather == "peer" else
                                       f", NCCL seam exchange between neighbours + send/recv gather of owned rows to rank 0 ({args.gather})")
                                      if world > 1 else ""),
                       "l2": "inputs larger than L2 (288 MB image, >1 GB activation arena per batch)",
                       "algorithmic_tflop_per_step": flops_image / 1e12},
            "pct_of_bf16_peak": {"sustained": flops_image / (ms / args.steps * 1e-3) / 1e12 / (pk["sustained"] * world),
                                 "burst": flops_image / (ms / args.steps * 1e-3) / 1e12 / (pk["burst"] * world)},
            "e2e": {"value": MP * e2e_steps / (e2e_ms * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "throughput_mode": thr, "parity": parity,
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
