"""torchrun --nproc-per-node N tools/dist_phases.py [cs ucs]: where the time of the sharded paths goes
(max over ranks, CUDA events / wall clock around barriers)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402
from nind_denoise_b200.tiler import _band, _nx, default_batch  # noqa: E402

cs, ucs = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (248, 224)
ol, W, H = 6, 6000, 4000
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()
img_host = torch.rand((3, H, W), generator=torch.Generator().manual_seed(1)).pin_memory()
img = img_host.to(dev)
n = nb.n_crops(W, H, cs, ucs, ol)
ranges = nb.shard_ranges(n, world)
cb, ce = ranges[rank]
batch = default_batch(ce - cb, cs, _nx(W, ucs, ol))
ext = nb.band_extents(W, H, cs, ucs, ol, ranges)
own = nb.owned_rows(ext, H)
shared = nb.SharedHostImage((3, H, W))
plain = torch.empty((3, H, W)).pin_memory() if rank == 0 else None


def timed(name, fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{name:46s} {ms.item():8.3f} ms  -> {24.0 / ms.item() * 1e3:7.0f} MP/s", flush=True)


def band_only():
    return _band(model, img, cs, ucs, ol, cb, ce, batch)


def band_exchange():
    band, y0, y1 = band_only()
    nb.exchange_seams(band, ext, own, rank)


def steps_only():
    from nind_denoise_b200.tiler import _peer_cache
    full = _peer_cache.setdefault(("dbg_full",), torch.empty((3, H, W), device=dev))
    for (a, b) in nb.plan_steps(model, W, H, cs, ucs, ol, cb, ce, batch):
        nb.tiled_step(model, img, full, cs, ucs, ol, cb, ce, a, b)


def peer_copy_only():
    from nind_denoise_b200.tiler import _peer_cache, _peer_gather, copy_planes
    pg = _peer_gather(H, W, 512, dev, None, 0)
    full = _peer_cache.setdefault(("dbg_full",), torch.empty((3, H, W), device=dev))
    o0, o1 = own[rank]
    copy_planes(pg.out()[:, o0:o1], full[:, o0:o1])


if rank == 0:
    print(f"world {world} cs {cs}: {n} crops, {ce - cb} per rank, batch {batch}; steps "
          f"{nb.plan_steps(model, W, H, cs, ucs, ol, cb, ce, batch)}", flush=True)
timed("band only (no communication)", band_only)
timed("steps only (forward + stitch per step)", steps_only)
timed("peer copy of the owned rows only", peer_copy_only)
timed("device path, peer mode", lambda: nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch, mode="peer"))
timed("band + seam exchange", band_exchange)
timed("device path, rows mode", lambda: nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch, mode="rows"))
timed("device path, bands mode", lambda: nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch, mode="bands"))
timed("host path, shared host image (pipelined)", lambda: nb.denoise_tiled_distributed_host(img_host, model, cs, ucs, ol, batch=batch, out=shared))
timed("host path, gather to rank 0 + one D2H", lambda: nb.denoise_tiled_distributed_host(img_host, model, cs, ucs, ol, batch=batch, out=plain))
ref = nb.denoise_tiled_distributed(img, model, cs, ucs, ol, batch=batch)
if rank == 0:
    print("shared vs device path max diff", float((shared.tensor.to(dev) - ref).abs().max()), flush=True)
shared.close()
dist.barrier()
dist.destroy_process_group()
