"""One GPU, one rank's share of an 8-GPU job through the host-range pipeline: step policy sweep.
python tools/host_range_tune.py [world] [cs ucs]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402
from nind_denoise_b200 import _capi  # noqa: E402
from nind_denoise_b200.tiler import _nx, default_batch, host_range  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cs, ucs = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (248, 224)
ol, W, H = 6, 6000, 4000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()
img_host = torch.rand((3, H, W)).pin_memory()
out_host = torch.empty((3, H, W)).pin_memory()
n = nb.n_crops(W, H, cs, ucs, ol)
ranges = nb.shard_ranges(n, world)
ext = nb.band_extents(W, H, cs, ucs, ol, ranges)
own = nb.owned_rows(ext, H)
lib = _capi.lib()
h = model.native_handle()


def run(r):
    cb, ce = ranges[r]
    o0, o1 = own[r]
    lo = min(o1, max([o0] + [ext[q][1] for q in range(r)]))
    b = default_batch(ce - cb, cs, _nx(W, ucs, ol))
    host_range(model, img_host, out_host, cs, ucs, ol, b, cb, ce, lo, o1)
    _capi.check(lib.nind_host_sync(h))


def timed(r, reps=8):
    run(r); run(r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        run(r)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


nx = _nx(W, ucs, ol)
per = ranges[0][1] - ranges[0][0]
print(f"world {world} cs {cs}: {per} crops/rank, nx {nx}")
for first, last in ((-1, -1), (0, 0), (per // 4, 0), (per // 4, per // 4), (per // 3, per // 3), (per // 6, per // 3),
                    (per // 8, per // 4), (nx // 2, nx), (nx // 2, nx // 2)):
    model.set_option("host_first", first)
    model.set_option("host_last", last)
    ts = [timed(r) for r in sorted({0, world // 2, world - 1})]
    print(f"first {first:3d} last {last:3d}: " + "  ".join(f"{t:6.3f} ms" for t in ts) + f"   max {max(ts):.3f}")
