"""Crop-size / overlap sweep (BASELINE configs[4]) on one GPU, device-resident and host-to-host:
python tools/sweep_cs.py > profiles/rNN_sweep_cs.log"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()
img_host = torch.rand((3, 4000, 6000), generator=torch.Generator().manual_seed(1)).pin_memory()
out_host = torch.empty_like(img_host).pin_memory()
img = img_host.to(dev)


def timed(fn, steps=4):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


print("cs ucs overlap crops | device-resident ms MP/s | host-to-host ms MP/s")
for cs, pads in ((120, (12,)), (248, (12, 24)), (504, (12, 24)), (1016, (12, 32))):
    for pad in pads:
        ucs = cs - 2 * pad
        for ol in (6, 32):
            if ol >= ucs // 2:
                continue
            n = nb.n_crops(6000, 4000, cs, ucs, ol)
            a = timed(lambda: nb.denoise_tiled(img, model, cs, ucs, ol))
            b = timed(lambda: nb.denoise_tiled_host(img_host, model, cs, ucs, ol, out=out_host))
            print(f"{cs:5d} {ucs:5d} {ol:3d} {n:6d} | {a:8.2f} {24000 / a:7.1f} | {b:8.2f} {24000 / b:7.1f}", flush=True)
