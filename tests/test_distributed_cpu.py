"""CPU, world_size 2 over gloo: crop sharding + band gather + assembly of the multi-GPU path, with the
per-rank band computed by the oracle (the GPU band kernel itself is covered by -m gpu tests)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nind_denoise_b200 as nb
from oracle import geometry as og

W, H, CS, UCS, OL = 101, 83, 40, 28, 4


def fake_model_np(c):
    x = torch.from_numpy(c)
    ramp = torch.linspace(0.5, 1.5, x.shape[-1]).view(1, 1, -1) * torch.linspace(1.25, 0.75, x.shape[-2]).view(1, -1, 1)
    return (x * ramp + 0.125).numpy()


def oracle_band(img, a, b):
    g = og.crop_grid(W, H, CS, UCS, OL)
    full = og.stitch(lambda i: fake_model_np(og.gather_crop(img.numpy(), g, i)), g, (a, b))
    t = og.crop_table(g)
    y0, y1 = int(t[a, 7]), min(H, int(t[b - 1, 7] + t[b - 1, 5] - t[b - 1, 3]))
    assert not full[:, :y0].any() and not full[:, y1:].any()
    return torch.from_numpy(full[:, y0:y1].copy()), y0, y1


def _worker(rank, world, port, q, shared=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    img = torch.from_numpy(np.random.default_rng(3).random((3, H, W), dtype=np.float32))
    if shared:  # seam exchange between neighbours + every rank writes its own rows into one shared host image
        sh = nb.SharedHostImage((3, H, W))
        sh.tensor.fill_(float("nan")) if rank == 0 else None
        dist.barrier()
        for _ in range(2):  # the buffer is reusable
            out = nb.denoise_tiled_distributed_host(img, None, CS, UCS, OL, out=sh, band_fn=oracle_band)
    else:
        out = nb.denoise_tiled_distributed(img, None, CS, UCS, OL, band_fn=oracle_band)
        out_b = nb.denoise_tiled_distributed(img, None, CS, UCS, OL, band_fn=oracle_band, mode="bands")
        if rank == 0:
            assert float((out - out_b).abs().max()) <= 1e-6
    if rank == 0:
        ref = og.denoise_tiled(img.numpy(), fake_model_np, CS, UCS, OL)
        q.put(float(np.abs(out.numpy() - ref).max()))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _run(world, shared=False):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, shared)) for r in range(world)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-6


def test_two_rank_gather_equals_single():
    _run(2)


def test_eight_ranks_with_an_empty_shard():
    # 20 crops over 8 ranks -> ceil = 3 per rank, rank 7 gets nothing
    assert nb.shard_ranges(nb.n_crops(W, H, CS, UCS, OL), 8)[-1] == (20, 20)
    _run(8)


def test_two_ranks_shared_host_image():
    _run(2, shared=True)


def test_eight_ranks_shared_host_image_with_an_empty_shard():
    _run(8, shared=True)


def test_owned_rows_partition_the_image():
    for world in (1, 2, 3, 5, 8, 19, 40):
        for (w, h, cs, ucs, ol) in ((101, 83, 40, 28, 4), (600, 400, 56, 40, 6), (64, 300, 72, 40, 2)):
            ranges = nb.shard_ranges(nb.n_crops(w, h, cs, ucs, ol), world)
            ext = nb.band_extents(w, h, cs, ucs, ol, ranges)
            own = nb.owned_rows(ext, h)
            rows = np.zeros(h, dtype=np.int32)
            for (o0, o1), (y0, y1) in zip(own, ext):
                assert o0 >= y0 and (o1 <= y1 or o1 == o0)   # a rank only owns rows of its own band
                rows[o0:o1] += 1
            assert (rows == 1).all()


def test_owned_rows_up_partition_and_seam_plan_is_consistent():
    """Ownership of the host-buffer entry (the EARLIER rank owns the grid row two ranges share): a partition of the
    image rows inside each owner's band; every send has exactly one matching receive; rows are only handed to
    earlier ranks, so a rank's hand-over is ready after its first step(s)."""
    for world in (1, 2, 3, 5, 8, 19, 40):
        for (w, h, cs, ucs, ol) in ((101, 83, 40, 28, 4), (600, 400, 56, 40, 6), (64, 300, 72, 40, 2)):
            ranges = nb.shard_ranges(nb.n_crops(w, h, cs, ucs, ol), world)
            ext = nb.band_extents(w, h, cs, ucs, ol, ranges)
            own = nb.owned_rows_up(ext, h)
            rows = np.zeros(h, dtype=np.int32)
            for (o0, o1), (y0, y1) in zip(own, ext):
                assert o1 == o0 or (o0 >= y0 and o1 <= y1)
                rows[o0:o1] += 1
            assert (rows == 1).all()
            plans = [nb.seam_plan(ext, own, r) for r in range(world)]
            sends = sorted((r, dst, a, b) for r, (s, _) in enumerate(plans) for dst, a, b in s)
            recvs = sorted((src, r, a, b) for r, (_, rc) in enumerate(plans) for src, a, b in rc)
            assert sends == recvs
            assert all(dst < src for src, dst, _, _ in sends)
            covered = np.zeros(h, dtype=np.int32)   # every band row is either owned or handed to its owner
            for r, (y0, y1) in enumerate(ext):
                mine = np.zeros(h, dtype=bool)
                mine[own[r][0]:own[r][1]] = True
                for _, a, b in plans[r][0]:
                    mine[a:b] = True
                assert mine[y0:y1].all()
