"""torchrun --nproc-per-node N tools/dist_phases_host.py [cs ucs]: per-rank phase break-down of the multi-GPU
host-buffer entry (denoise_tiled_distributed_host with a SharedHostImage): enqueue, pipeline (H2D | forward | stitch
| D2H of the rows no earlier rank touches), seam exchange, seam-row D2H, closing synchronisation.  Phases are
separated by stream synchronisations (which serialise what would otherwise overlap: read the table as an upper
bound per phase, next to the un-instrumented total printed last)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402
from nind_denoise_b200 import _capi  # noqa: E402
from nind_denoise_b200.tiler import _nx, default_batch, host_range  # noqa: E402

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
n = nb.n_crops(W, H, cs, ucs, ol)
ranges = nb.shard_ranges(n, world)
cb, ce = ranges[rank]
batch = default_batch(ce - cb, cs, _nx(W, ucs, ol))
ext = nb.band_extents(W, H, cs, ucs, ol, ranges)
own = nb.owned_rows(ext, H)
shared = nb.SharedHostImage((3, H, W))
y0, y1 = ext[rank]
o0, o1 = own[rank]
lo = min(o1, max([o0] + [ext[r][1] for r in range(rank) if ext[r][1] > ext[r][0]]))
lib = _capi.lib()


def instrumented():
    t = [time.perf_counter()]
    full = host_range(model, img_host, shared.tensor, cs, ucs, ol, batch, cb, ce, lo, o1)
    band = full[:, y0:y1, :]
    t.append(time.perf_counter())                       # enqueue (CPU)
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())                       # compute + stitch (joined on the torch stream)
    _capi.check(lib.nind_host_sync(model.native_handle()))
    t.append(time.perf_counter())                       # pipelined D2H tail
    nb.exchange_seams(band, ext, own, rank)
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())                       # seam exchange
    if lo > o0:
        for c in range(3):
            shared.tensor[c, o0:lo].copy_(band[c, o0 - y0:lo - y0], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    t.append(time.perf_counter())                       # seam rows D2H
    dist.barrier()
    t.append(time.perf_counter())                       # closing barrier
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]


for _ in range(3):
    instrumented()
torch.cuda.synchronize(); dist.barrier()
acc = None
reps = 6
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier()
    p = instrumented()
    acc = p if acc is None else [a + b for a, b in zip(acc, p)]
mine = torch.tensor([a / reps for a in acc] + [float(ce - cb), float(lo - o0), float(o1 - lo)], device=dev)
allp = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allp, mine)
if rank == 0:
    print(f"world {world} cs {cs}: {n} crops, batch {batch}; ms per phase (each phase ends with a synchronisation)")
    print("rank crops seam_rows piped_rows | enqueue  compute  d2h_tail  exchange  seam_d2h  barrier | sum")
    for r, v in enumerate(allp):
        v = v.tolist()
        print(f"{r:4d} {int(v[6]):5d} {int(v[7]):9d} {int(v[8]):10d} | " + "  ".join(f"{x:7.3f}" for x in v[:6]) + f" | {sum(v[:6]):6.3f}")


def plain():
    nb.denoise_tiled_distributed_host(img_host, model, cs, ucs, ol, batch=batch, out=shared)


for _ in range(2):
    plain()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    plain()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"un-instrumented total: {ms.item():.3f} ms per image -> {24.0 / ms.item() * 1e3:.0f} MP/s")
shared.close()
dist.barrier()
dist.destroy_process_group()
