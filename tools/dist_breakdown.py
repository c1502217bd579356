"""Single-GPU break-down of what one rank of an N-GPU job does: band compute for 1/N of the crops
(small-batch efficiency), pinned H2D / D2H rates.  python tools/dist_breakdown.py [cs ucs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402
from nind_denoise_b200.tiler import _band, _nx, default_batch, rows_needed  # noqa: E402

cs, ucs = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (248, 224)
ol, W, H = 6, 6000, 4000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()
img_host = torch.rand((3, H, W)).pin_memory()
img = img_host.to(dev)
n = nb.n_crops(W, H, cs, ucs, ol)


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for world in (1, 2, 4, 8):
    ranges = nb.shard_ranges(n, world)
    for r in sorted({0, world // 2, world - 1}):
        cb, ce = ranges[r]
        b = default_batch(ce - cb, cs, _nx(W, ucs, ol))
        ms = timed(lambda: _band(model, img, cs, ucs, ol, cb, ce, b))
        r0, r1 = rows_needed(W, H, cs, ucs, ol, cb, ce)
        print(f"world {world} rank {r}: crops {ce - cb} batch {b}: band {ms:.3f} ms = {ms / (ce - cb) * 1e3:.1f} us/crop "
              f"(ideal N-GPU rate {24.0 / ms * 1e3 * (ce - cb) / n * world:.0f} MP/s); rows needed {r1 - r0}")
out_host = torch.empty_like(img_host).pin_memory()
for rows in (4000, 2000, 1000, 500):
    d = torch.empty((3, rows, W), device=dev)
    h = img_host[:, :rows].contiguous().pin_memory()
    ms_u = timed(lambda: d.copy_(h, non_blocking=True))
    ms_d = timed(lambda: h.copy_(d, non_blocking=True))
    gb = d.numel() * 4 / 1e9
    print(f"pinned {gb * 1e3:.0f} MB: H2D {ms_u:.3f} ms = {gb / ms_u * 1e3:.1f} GB/s, D2H {ms_d:.3f} ms = {gb / ms_d * 1e3:.1f} GB/s")
