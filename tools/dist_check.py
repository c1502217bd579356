"""torchrun --nproc-per-node N tools/dist_check.py : the sharded multi-GPU path against the single-GPU path."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = nb.UtNet().to(dev).eval()
g = torch.Generator().manual_seed(5)
img = torch.rand((3, 1100, 1500), generator=g).to(dev)
for cs, ucs, ol in ((248, 224, 6), (120, 96, 6)):
    out = nb.denoise_tiled_distributed(img, model, cs, ucs, ol)
    if dist.get_rank() == 0:
        ref = nb.denoise_tiled(img, model, cs, ucs, ol)
        err = float((out - ref).abs().max())
        print(f"world {dist.get_world_size()} cs {cs}: max |distributed - single| = {err:.3e}")
        assert err <= 1e-6
    outh = nb.denoise_tiled_distributed_host(img.cpu().pin_memory(), model, cs, ucs, ol)
    if dist.get_rank() == 0:
        errh = float((outh.to(dev) - ref).abs().max())
        print(f"world {dist.get_world_size()} cs {cs}: host entry max diff = {errh:.3e}")
        assert errh <= 1e-6
    sh = nb.SharedHostImage(tuple(img.shape))
    for _ in range(2):
        outs = nb.denoise_tiled_distributed_host(img.cpu().pin_memory(), model, cs, ucs, ol, out=sh)
    if dist.get_rank() == 0:
        errs = float((outs.to(dev) - ref).abs().max())
        print(f"world {dist.get_world_size()} cs {cs}: shared host image max diff = {errs:.3e} (pinned {sh.pinned})")
        assert errs <= 1e-6
    sh.close()
dist.barrier()
dist.destroy_process_group()
