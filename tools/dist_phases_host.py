"""torchrun --nproc-per-node N tools/dist_phases_host.py [cs ucs]: per-rank timeline of the multi-GPU host-buffer
entry (denoise_tiled_distributed_host with a SharedHostImage), from the call's OWN blocking points (its `phases`
marks; no extra synchronisation is added): enqueue of the H2D | forward | stitch | D2H pipeline, hand-over of the
rows an earlier rank owns (peer DMA after the first step), arrival of the later rank's rows, all own rows landed in
the shared host image (pipeline + add + tail D2H), every rank arrived (rank 0 only waits)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402
from nind_denoise_b200.tiler import _nx, default_batch  # noqa: E402

cs, ucs = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (248, 224)
ol, W, H = 6, 6000, 4000
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("NIND_NO_NUMA_BIND") is None:
    print(f"rank {local}: bound to {len(nb.bind_host_to_gpu(local) or [])} CPUs of the GPU's NUMA node", flush=True)
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
own = nb.owned_rows_up(ext, H)
sends, recvs = nb.seam_plan(ext, own, rank)
shared = nb.SharedHostImage((3, H, W))
NAMES = ["enqueued", "handed_over", "received", "rows_landed", "all_arrived"]


def run(phases=None):
    nb.denoise_tiled_distributed_host(img_host, model, cs, ucs, ol, batch=batch, out=shared, phases=phases)


for _ in range(3):
    run()
reps = 8
acc = [0.0] * len(NAMES)
for _ in range(reps):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ph = []
    run(ph)
    t0 = ph[0][1]
    d = dict(ph)
    last = t0
    for i, k in enumerate(NAMES):
        last = d.get(k, last)
        acc[i] += 1e3 * (last - t0)
mine = torch.tensor([a / reps for a in acc] + [float(ce - cb), float(sum(b - a for _, a, b in sends)),
                                               float(sum(b - a for _, a, b in recvs)), float(own[rank][1] - own[rank][0])],
                    device=dev)
allp = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allp, mine)
if rank == 0:
    print(f"world {world} cs {cs}: {n} crops, batch {batch}; ms since the start of the call (host clock, no added syncs)")
    print("rank crops rows_sent rows_recv rows_owned | enqueued  handed_over  received  rows_landed  all_arrived")
    for r, v in enumerate(allp):
        v = v.tolist()
        print(f"{r:4d} {int(v[5]):5d} {int(v[6]):9d} {int(v[7]):9d} {int(v[8]):10d} | " + "  ".join(f"{x:10.3f}" for x in v[:5]))

for _ in range(2):
    run()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    run()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"back-to-back total: {ms.item():.3f} ms per image -> {24.0 / ms.item() * 1e3:.0f} MP/s")
shared.close()
dist.barrier()
dist.destroy_process_group()
