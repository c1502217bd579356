"""Same-process A/B of plan options (batch size, n_tile_deep, cta_group) — avoids box-to-box variance."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb

dev = torch.device("cuda:0")
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
img = torch.rand((3, 4000, 6000), generator=g).to(dev)

def run(cs, batch, opts, steps=6):
    torch.manual_seed(0)
    m = nb.UtNet().to(dev).eval()
    for k, v in opts.items():
        m.set_option(k, v)
    ucs = cs - 24
    for _ in range(2):
        nb.denoise_tiled(img, m, cs, ucs, 6, batch=batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        nb.denoise_tiled(img, m, cs, ucs, 6, batch=batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"cs {cs} batch {batch:4d} {opts}: {ms:7.2f} ms  {24000/ms:7.1f} MP/s", flush=True)
    del m
    torch.cuda.empty_cache()

for rep in range(2):
    for batch in (112, 168, 224, 280):
        run(248, batch, {})
    for batch in (26, 39, 52, 65):
        run(504, batch, {})
    run(504, 39, {"n_tile_deep": 128})
    run(248, 168, {"n_tile_deep": 128})
    run(504, 39, {"cta_group": 1})
