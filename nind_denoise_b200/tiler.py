"""Tiled denoising of a whole image — the inner loop of the reference's denoise_image.py
(/root/reference/src/nind_denoise/denoise_image.py:231-267) behind one call.

``denoise_tiled(img, model, cs, ucs, ol)`` keeps the reference's semantics exactly: the same crop
grid (``OneImageDS.__init__``), mirror-padded gather (``__getitem__``), trim by ``usefuldim``,
``make_seamless_edges`` halving and raster-order overlap-add.  All of it runs on the GPU through
``nind_tiled_denoise`` (include/nind_b200.h); crops are processed ``batch`` at a time.

Multi-GPU (one process per GPU, torch.distributed): crops are split into contiguous raster ranges,
every rank stitches its own row band, and the bands are gathered to rank 0 — the only collective.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _capi

# reference defaults, denoise_image.py:40-42
CS_UNET, UCS_UNET = 440, 320
CS_UTNET, UCS_UTNET = 504, 480
DEFAULT_OVERLAP = 6


def crop_table(width: int, height: int, cs: int, ucs: int, ol: int):
    """int32 [n_crops, 8]: x0, y0, usefuldim (x_lo, y_lo, x_hi, y_hi), usefulstart (x, y)."""
    return _capi.crop_table(width, height, cs, ucs, ol)


def _nx(width: int, ucs: int, ol: int) -> int:
    return math.ceil((width - ucs) / (ucs - ol)) + 1


def n_crops(width: int, height: int, cs: int, ucs: int, ol: int) -> int:
    stride = ucs - ol
    return (math.ceil((width - ucs) / stride) + 1) * (math.ceil((height - ucs) / stride) + 1)


def default_batch(n: int, cs: int, nx: int = 0) -> int:
    """Crops per forward.  Large enough that the deep (small-map) layers fill 148 SMs for several waves
    (B200 sweep: cs 248 -> 451/473/484/489 MP/s at 16/28/56/84 crops; cs 504 -> 524/528/534 at 13/26/39),
    rounded to whole grid rows so the host pipeline (H2D | compute | D2H) steps row by row."""
    target = int(max(4, min(160, round(9.5e6 / float(cs * cs)))))
    if nx and target >= nx:
        target = max(1, round(target / nx)) * nx
    return max(1, min(n, target))


def shard_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous raster ranges of crop indices, ceil(n/world) per rank; trailing ranks
    may be empty."""
    per = -(-n // world)
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def _band(model, img: torch.Tensor, cs, ucs, ol, crop_begin, crop_end, batch) -> Tuple[torch.Tensor, int, int]:
    """Rows [y0, y1) touched by crops [crop_begin, crop_end) -> ([3, y1-y0, W] fp32 on img.device, y0, y1)."""
    if not img.is_cuda:
        raise RuntimeError("denoise_tiled (nind_denoise_b200): image must be a CUDA tensor; there is no CPU path")
    _, H, W = img.shape
    h = model.native_handle()
    y0, y1 = _capi.band_rows(W, H, cs, ucs, ol, crop_begin, crop_end)
    out = torch.empty((3, y1 - y0, W), dtype=torch.float32, device=img.device)
    by0, by1 = C.c_int(), C.c_int()
    with torch.cuda.device(img.device):
        stream = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().nind_tiled_denoise(h, img.data_ptr(), out.data_ptr(), H, W, cs, ucs, ol, crop_begin,
                                                   crop_end, batch, C.byref(by0), C.byref(by1), C.c_void_p(stream)))
    assert (by0.value, by1.value) == (y0, y1)
    return out, y0, y1


def denoise_tiled(img: torch.Tensor, model, cs: Optional[int] = None, ucs: Optional[int] = None,
                  ol: int = DEFAULT_OVERLAP, batch: Optional[int] = None) -> torch.Tensor:
    """[3,H,W] fp32 CUDA image -> [3,H,W] fp32 denoised image on the same device (no clamp, as the
    reference's '.tiff' path, pt_helpers.py:30-32)."""
    if img.dim() != 3 or img.shape[0] != 3:
        raise ValueError(f"expected a [3,H,W] image, got {tuple(img.shape)}")
    if cs is None or ucs is None:  # autodetect_network_cs_ucs, denoise_image.py:59-79
        cs, ucs = (CS_UTNET, UCS_UTNET) if type(model).__name__ == "UtNet" else (CS_UNET, UCS_UNET)
    img = img.detach().float().contiguous()
    n = n_crops(img.shape[2], img.shape[1], cs, ucs, ol)
    if batch is None:
        batch = default_batch(n, cs, _nx(img.shape[2], ucs, ol))
    out, y0, y1 = _band(model, img, cs, ucs, ol, 0, n, batch)
    assert y0 == 0 and y1 == img.shape[1]
    return out


def pad_whole_image(img: torch.Tensor, pad: int) -> torch.Tensor:
    """The reference's whole-image input (denoise_image.py:110-126): the image centred in a zero canvas
    ``pad`` larger on every side, the four sides filled with the edge-inclusive mirror of the image, the
    corners left at zero.  (The reference allocates the canvas with width and height swapped, :113, so it
    only runs on square images; this is the layout it builds for those.)"""
    _, H, W = img.shape
    if pad < 0 or pad > min(H, W):
        raise ValueError(f"pad must be in [0, {min(H, W)}], got {pad}")
    ret = torch.zeros((3, H + 2 * pad, W + 2 * pad), dtype=img.dtype, device=img.device)
    ret[:, pad:H + pad, pad:W + pad] = img
    if pad:
        ret[:, pad:H + pad, :pad] = img[:, :, :pad].flip(2)
        ret[:, pad:H + pad, W + pad:] = img[:, :, W - pad:].flip(2)
        ret[:, :pad, pad:W + pad] = img[:, :pad, :].flip(1)
        ret[:, H + pad:, pad:W + pad] = img[:, H - pad:, :].flip(1)
    return ret


def denoise_whole_image(img: torch.Tensor, model, pad: int = 0) -> torch.Tensor:
    """``--whole_image`` mode of the reference script (denoise_image.py:91-97,110-128,255-256): ONE forward
    over the mirror-padded image, trimmed back to [3,H,W].  For small images; the padded size must be one
    the network accepts (UtNet: 16a+56 — the reference fails inside torch.cat otherwise, this raises)."""
    if img.dim() != 3 or img.shape[0] != 3:
        raise ValueError(f"expected a [3,H,W] image, got {tuple(img.shape)}")
    if not img.is_cuda:
        raise RuntimeError("denoise_whole_image (nind_denoise_b200): image must be a CUDA tensor; there is no CPU path")
    _, H, W = img.shape
    x = pad_whole_image(img.detach().float(), int(pad or 0))
    y = model(x.unsqueeze(0))[0]
    pad = int(pad or 0)
    return y[:, pad:H + pad, pad:W + pad].contiguous()


def denoise_tiled_host(img_host: torch.Tensor, model, cs: int, ucs: int, ol: int = DEFAULT_OVERLAP,
                       batch: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Host-buffer variant (the call the reference's script would make): H2D + all crops + D2H inside
    ``nind_tiled_denoise_host``.  ``img_host`` / ``out`` are CPU tensors (pinned for full PCIe speed)."""
    if img_host.is_cuda:
        raise ValueError("denoise_tiled_host expects a CPU tensor")
    img_host = img_host.detach().float().contiguous()
    _, H, W = img_host.shape
    if out is None:
        out = torch.empty_like(img_host)
    if batch is None:
        batch = default_batch(n_crops(W, H, cs, ucs, ol), cs, _nx(W, ucs, ol))
    h = model.native_handle()
    with torch.cuda.device(model._device):
        _capi.check(_capi.lib().nind_tiled_denoise_host(h, img_host.data_ptr(), out.data_ptr(), H, W, cs, ucs, ol,
                                                        batch))
    return out


def denoise_images_host(imgs_host: Sequence[torch.Tensor], model, cs: int, ucs: int, ol: int = DEFAULT_OVERLAP,
                        batch: Optional[int] = None, outs: Optional[Sequence[torch.Tensor]] = None):
    """Throughput mode on one GPU: a sequence of CPU images (pinned for full PCIe speed) is pushed through
    ``nind_tiled_denoise_host_async`` back to back — image k+1's upload overlaps image k's compute and
    download — and synchronised once.  Returns the list of CPU outputs."""
    imgs = [im.detach().float().contiguous() for im in imgs_host]
    if outs is None:
        outs = [torch.empty_like(im).pin_memory() if im.is_pinned() else torch.empty_like(im) for im in imgs]
    h = model.native_handle()
    lib = _capi.lib()
    with torch.cuda.device(model._device):
        for im, out in zip(imgs, outs):
            _, H, W = im.shape
            b = batch or default_batch(n_crops(W, H, cs, ucs, ol), cs, _nx(W, ucs, ol))
            _capi.check(lib.nind_tiled_denoise_host_async(h, im.data_ptr(), out.data_ptr(), H, W, cs, ucs, ol, b))
        _capi.check(lib.nind_host_sync(h))
    return list(outs)


# ------------------------------------------------------------------------------ multi-GPU
def assemble_bands(bands: Sequence[Tuple[Optional[torch.Tensor], int, int]], height: int, width: int,
                   device=None) -> torch.Tensor:
    """Sum per-rank row bands into the [3,H,W] image (rank order = raster order, so a pixel's
    contributions are added in increasing crop index between ranks)."""
    first = next(b for b, _, _ in bands if b is not None)
    out = torch.zeros((3, height, width), dtype=torch.float32, device=device or first.device)
    for band, y0, y1 in bands:
        if band is not None and y1 > y0:
            out[:, y0:y1, :] += band.to(out.device)
    return out


def denoise_tiled_distributed(img: torch.Tensor, model, cs: int, ucs: int, ol: int = DEFAULT_OVERLAP,
                              batch: Optional[int] = None, group=None, dst: int = 0,
                              band_fn: Optional[Callable] = None, mode: str = "rows") -> Optional[torch.Tensor]:
    """Every rank holds the same ``img`` (read-only) and a replica of ``model``; rank ``dst`` returns the
    stitched image, the others return None.

    ``mode="rows"`` (default): neighbours first exchange the seam rows they share (``exchange_seams``),
    then every rank sends only the rows it owns and ``dst`` receives them straight into the output image —
    each output byte crosses NVLink once and nothing is summed on ``dst``.
    ``mode="bands"``: whole overlapping bands are gathered and summed on ``dst`` (kept for comparison).

    ``band_fn(img, crop_begin, crop_end) -> (band, y0, y1)`` replaces the GPU band computation in CPU
    (gloo) tests of the sharding/gather logic."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    _, H, W = img.shape
    n = n_crops(W, H, cs, ucs, ol)
    ranges = shard_ranges(n, world)
    cb, ce = ranges[rank]
    if band_fn is None:
        if batch is None:
            batch = default_batch(max(1, ce - cb), cs, _nx(W, ucs, ol))
        band_fn = lambda im, a, b: _band(model, im, cs, ucs, ol, a, b, batch)
    # band extents are pure geometry: every rank can compute everybody's (no metadata exchange)
    extents = band_extents(W, H, cs, ucs, ol, ranges)
    band = None
    if ce > cb:
        band, y0, y1 = band_fn(img, cb, ce)
        assert (y0, y1) == extents[rank]
    if mode == "rows":
        own = owned_rows(extents, H)
        exchange_seams(band, extents, own, rank, group)
        o0, o1 = own[rank]
        ops = []
        if rank == dst:
            out = torch.empty((3, H, W), dtype=torch.float32, device=img.device)
            for r in range(world):
                a, b = own[r]
                if b <= a:
                    continue
                if r == rank:
                    out[:, a:b, :].copy_(band[:, a - y0:b - y0, :])
                else:  # one receive per colour plane: out[c, a:b] is contiguous
                    ops += [dist.P2POp(dist.irecv, out[c, a:b, :], r, group) for c in range(3)]
        elif o1 > o0:
            ops = [dist.P2POp(dist.isend, band[c, o0 - y0:o1 - y0, :], dst, group) for c in range(3)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out if rank == dst else None
    if mode != "bands":
        raise ValueError(f"unknown mode {mode!r}")
    if rank == dst:
        # post every receive at once (one batched NCCL group): the bands arrive concurrently through NVSwitch
        bands, ops = [], []
        for r in range(world):
            y0, y1 = extents[r]
            if y1 <= y0:
                bands.append((None, 0, 0))
            elif r == rank:
                bands.append((band, y0, y1))
            else:
                buf = torch.empty((3, y1 - y0, W), dtype=torch.float32, device=img.device)
                ops.append(dist.P2POp(dist.irecv, buf, r, group))
                bands.append((buf, y0, y1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return assemble_bands(bands, H, W, device=img.device)
    if band is not None:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, band.contiguous(), dst, group)]):
            req.wait()
    return None


def rows_needed(width: int, height: int, cs: int, ucs: int, ol: int, crop_begin: int, crop_end: int) -> Tuple[int, int]:
    """Image rows [r0, r1) that crops [crop_begin, crop_end) read, INCLUDING the rows the tiler's mirror
    padding reflects to when a crop window sticks out of the image (at the bottom edge the window can
    overshoot by most of a crop, so the mirrored rows lie above the window's own first row)."""
    t = crop_table(width, height, cs, ucs, ol)
    y_first, y_last_end = int(t[crop_begin, 1]), int(t[crop_end - 1, 1]) + cs
    r0, r1 = max(0, y_first), min(height, y_last_end)
    if y_last_end > height:   # rows height-1 ... 2*height - y_last_end are mirrored in
        r0 = min(r0, max(0, 2 * height - y_last_end))
    if y_first < 0:           # rows 0 ... -y_first-1 are mirrored in
        r1 = max(r1, min(height, -y_first))
    return r0, r1


def band_extents(width: int, height: int, cs: int, ucs: int, ol: int, ranges) -> List[Tuple[int, int]]:
    """Output rows [y0, y1) each rank's crop range touches ((0, 0) for an empty range) — pure geometry,
    every rank computes everybody's."""
    t = crop_table(width, height, cs, ucs, ol)
    ext = []
    for a, b in ranges:
        if b > a:
            ext.append((int(t[a, 7]), min(height, int(t[b - 1, 7]) + int(t[b - 1, 5] - t[b - 1, 3]))))
        else:
            ext.append((0, 0))
    return ext


def owned_rows(extents: Sequence[Tuple[int, int]], height: int) -> List[Tuple[int, int]]:
    """Disjoint row ownership for the scatter-free output: a non-empty rank owns the rows from its band's
    first row up to the next non-empty band's first row (the last one up to ``height``).  Rows of band r
    that lie in a later rank's range are that rank's to finish (see ``exchange_seams``)."""
    own = [(0, 0)] * len(extents)
    live = [r for r, (y0, y1) in enumerate(extents) if y1 > y0]
    for i, r in enumerate(live):
        end = extents[live[i + 1]][0] if i + 1 < len(live) else height
        own[r] = (extents[r][0], max(extents[r][0], end))
    return own


def exchange_seams(band: Optional[torch.Tensor], extents, own, rank: int, group=None) -> None:
    """Neighbour exchange that completes every rank's owned rows in place: rank r sends the rows of its
    band that a later rank owns (the grid row the two ranges share, or just the ``ol`` seam rows) and adds
    what earlier ranks send for its own rows, in rank (= raster) order.  One batched P2P group."""
    import torch.distributed as dist

    if band is None:
        return
    y0, y1 = extents[rank]
    ops, recvs, keep = [], [], []
    for s in range(rank + 1, len(extents)):
        a, b = max(y0, own[s][0]), min(y1, own[s][1])
        if b > a:
            keep.append(band[:, a - y0:b - y0, :].contiguous())
            ops.append(dist.P2POp(dist.isend, keep[-1], s, group))
    for r in range(rank):
        a, b = max(extents[r][0], own[rank][0]), min(extents[r][1], own[rank][1])
        if b > a:
            buf = torch.empty((3, b - a, band.shape[2]), dtype=band.dtype, device=band.device)
            ops.append(dist.P2POp(dist.irecv, buf, r, group))
            recvs.append((buf, a, b))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, a, b in recvs:
        band[:, a - y0:b - y0, :] += buf


class _DevicePtr:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def host_range(model, img_host: torch.Tensor, out_host: torch.Tensor, cs: int, ucs: int, ol: int, batch: int,
               crop_begin: int, crop_end: int, d2h_y0: int, d2h_y1: int) -> torch.Tensor:
    """``nind_tiled_denoise_host_range`` + ``nind_host_join``: enqueue one rank's crops on the library's
    H2D | compute | D2H pipeline (rows [d2h_y0, d2h_y1) go to ``out_host`` as they complete) and return the
    device image [3,H,W] its band is stitched into, ordered on the current torch stream."""
    if img_host.is_cuda or out_host.is_cuda or img_host.dtype != torch.float32 or not img_host.is_contiguous():
        raise ValueError("host_range expects contiguous fp32 CPU tensors")
    _, H, W = img_host.shape
    h = model.native_handle()
    lib = _capi.lib()
    d_out = C.c_void_p()
    with torch.cuda.device(model._device):
        _capi.check(lib.nind_tiled_denoise_host_range(h, img_host.data_ptr(), out_host.data_ptr(), H, W, cs, ucs, ol,
                                                      batch, crop_begin, crop_end, d2h_y0, d2h_y1, C.byref(d_out)))
        _capi.check(lib.nind_host_join(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return torch.as_tensor(_DevicePtr(d_out.value, (3, H, W)), device=model._device)


class SharedHostImage:
    """A [3,H,W] fp32 host image in ONE shared-memory segment mapped by every rank of ``group`` (ranks of a
    node), page-locked on each rank: the multi-GPU host entry lets every GPU copy the rows it owns straight
    into it over its own PCIe link instead of funnelling the whole image through rank ``src``'s.
    Collective constructor.  ``tensor`` is the CPU view (all ranks see the same bytes)."""

    def __init__(self, shape, group=None, src: int = 0, pin: bool = True):
        import os

        import torch.distributed as dist

        self.shape = tuple(int(v) for v in shape)
        numel = math.prod(self.shape)
        self._fd = None
        rank = dist.get_rank(group)
        path = [None]
        if rank == src:
            # anonymous memory file, reachable by the other ranks through /proc (no /dev/shm size limit)
            self._fd = os.memfd_create("nind_b200_out")
            os.ftruncate(self._fd, numel * 4)
            path[0] = f"/proc/{os.getpid()}/fd/{self._fd}"
        dist.broadcast_object_list(path, src=dist.get_global_rank(group, src) if group is not None else src,
                                   group=group)
        self.tensor = torch.from_file(path[0], shared=True, size=numel, dtype=torch.float32).view(self.shape)
        self.pinned = False
        if pin and torch.cuda.is_available():
            _capi.check(_capi.lib().nind_host_register(self.tensor.data_ptr(), numel * 4))
            self.pinned = True
        dist.barrier(group)  # everyone has mapped it; the creator can drop its descriptor
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None

    def close(self):
        if self.pinned:
            _capi.check(_capi.lib().nind_host_unregister(self.tensor.data_ptr()))
            self.pinned = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def denoise_tiled_distributed_host(img_host: torch.Tensor, model, cs: int, ucs: int, ol: int = DEFAULT_OVERLAP,
                                   batch: Optional[int] = None, group=None, dst: int = 0, out=None,
                                   band_fn: Optional[Callable] = None) -> Optional[torch.Tensor]:
    """Multi-GPU host-buffer entry: every rank holds the same CPU image (pinned for full PCIe speed),
    uploads only the rows its crop range reads and stitches its band on its GPU.

    * ``out`` a ``SharedHostImage``: neighbours exchange the seam rows over NCCL, then every rank copies
      the rows it owns straight into the shared host image — N PCIe links in parallel, no gather.
    * ``out`` a plain CPU tensor or None: bands are gathered to rank ``dst`` over NCCL and copied to host
      there (one PCIe link).

    Rank ``dst`` returns the CPU image, the others None.  ``band_fn(img, crop_begin, crop_end)`` stands in
    for the GPU band computation in the CPU (gloo) tests."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _, H, W = img_host.shape
    n = n_crops(W, H, cs, ucs, ol)
    ranges = shard_ranges(n, world)
    cb, ce = ranges[rank]
    if isinstance(out, SharedHostImage):
        if out.shape != (3, H, W):
            raise ValueError(f"shared output is {out.shape}, image is {(3, H, W)}")
        extents = band_extents(W, H, cs, ucs, ol, ranges)
        own = owned_rows(extents, H)
        y0, y1 = extents[rank]
        o0, o1 = own[rank]
        # own rows that earlier ranks' bands also touch are only final after the seam exchange
        lo = min(o1, max([o0] + [extents[r][1] for r in range(rank) if extents[r][1] > extents[r][0]]))
        band = full = None
        if ce > cb:
            if band_fn is None:
                if batch is None:
                    batch = default_batch(ce - cb, cs, _nx(W, ucs, ol))
                # H2D | forward | stitch | D2H of rows [lo, o1) pipelined inside the library; the stream we
                # continue on is ordered after its compute stream
                full = host_range(model, img_host, out.tensor, cs, ucs, ol, batch, cb, ce, lo, o1)
                band = full[:, y0:y1, :]
            else:
                band, by0, by1 = band_fn(img_host, cb, ce)
                assert (by0, by1) == (y0, y1)
                out.tensor[:, lo:o1, :].copy_(band[:, lo - y0:o1 - y0, :])
        exchange_seams(band, extents, own, rank, group)
        if lo > o0:
            for c in range(3):
                out.tensor[c, o0:lo].copy_(band[c, o0 - y0:lo - y0], non_blocking=True)
        if band_fn is None and ce > cb:
            _capi.check(_capi.lib().nind_host_sync(model.native_handle()))
            torch.cuda.current_stream(model._device).synchronize()
        dist.barrier(group)  # every rank's rows have landed in the shared image
        return out.tensor if rank == dst else None
    if band_fn is None:
        dev = model._device if getattr(model, "_handle", None) else next(model.parameters()).device
        d_img = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        if ce > cb:
            r0, r1 = rows_needed(W, H, cs, ucs, ol, cb, ce)
            for c in range(3):  # per plane: contiguous pinned source -> true async DMA (a strided CPU view is staged)
                d_img[c, r0:r1].copy_(img_host[c, r0:r1], non_blocking=True)
    else:
        dev, d_img = img_host.device, img_host
    res = denoise_tiled_distributed(d_img, model, cs, ucs, ol, batch=batch, group=group, dst=dst, band_fn=band_fn)
    if rank != dst:
        return None
    if out is None:
        out = torch.empty((3, H, W), dtype=torch.float32, pin_memory=dev.type == "cuda")
    out.copy_(res, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    return out


# ------------------------------------------------------------------------------ geometry ops
def gather_crops(model, img: torch.Tensor, cs: int, ucs: int, ol: int, crop_begin: int = 0,
                 crop_end: Optional[int] = None) -> torch.Tensor:
    """``OneImageDS.__getitem__`` for a range of crops: [3,H,W] CUDA image -> [n,3,cs,cs] fp32
    (bit-exact copies with the reference's mirror padding, denoise_image.py:129-174)."""
    _, H, W = img.shape
    if crop_end is None:
        crop_end = n_crops(W, H, cs, ucs, ol)
    img = img.detach().float().contiguous()
    out = torch.empty((crop_end - crop_begin, 3, cs, cs), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        stream = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().nind_gather_crops(model.native_handle(), img.data_ptr(), H, W, cs, ucs, ol, crop_begin,
                                                  crop_end, out.data_ptr(), C.c_void_p(stream)))
    return out


def stitch_crops(crops: torch.Tensor, height: int, width: int, cs: int, ucs: int, ol: int, crop_begin: int = 0,
                 crop_end: Optional[int] = None) -> Tuple[torch.Tensor, int, int]:
    """Trim + seam halving + overlap-add (denoise_image.py:204-213,250-267) of network outputs
    [n,3,cs,cs] -> (band [3, y1-y0, W], y0, y1)."""
    if crop_end is None:
        crop_end = crop_begin + crops.shape[0]
    assert crops.is_cuda and crops.shape[0] == crop_end - crop_begin and tuple(crops.shape[1:]) == (3, cs, cs)
    crops = crops.detach().float().contiguous()
    y0, y1 = _capi.band_rows(width, height, cs, ucs, ol, crop_begin, crop_end)
    out = torch.empty((3, y1 - y0, width), dtype=torch.float32, device=crops.device)
    by0, by1 = C.c_int(), C.c_int()
    with torch.cuda.device(crops.device):
        stream = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().nind_stitch_crops(crops.data_ptr(), height, width, cs, ucs, ol, crop_begin, crop_end,
                                                  out.data_ptr(), C.byref(by0), C.byref(by1), C.c_void_p(stream)))
    return out, y0, y1


class OneImageDS:
    """Mirror of the reference dataset class (denoise_image.py:81-177) over an in-memory CUDA image:
    ``ds[i]`` -> (crop [3,cs,cs], usefuldim IntTensor[4], usefulstart IntTensor[2]).  Tiled mode only."""

    def __init__(self, inimg, cs, ucs, ol, whole_image=False, pad=None, model=None):
        if whole_image:
            raise NotImplementedError("whole_image mode is not part of the tiled hot path (SURVEY §8f-3)")
        if not torch.is_tensor(inimg) or not inimg.is_cuda:
            raise RuntimeError("OneImageDS (nind_denoise_b200) takes a [3,H,W] CUDA tensor")
        self.inimg = inimg.detach().float().contiguous()
        self.height, self.width = self.inimg.shape[1], self.inimg.shape[2]
        self.cs, self.ucs, self.ol = cs, ucs, ol
        self.pad = int((cs - ucs) / 2)
        self.iperhl = math.ceil((self.width - ucs) / (ucs - ol))
        self.table = crop_table(self.width, self.height, cs, ucs, ol)
        self.size = self.table.shape[0]
        self._model = model

    def __len__(self):
        return self.size

    def __getitem__(self, i):
        if not 0 <= i < self.size:
            raise IndexError(i)
        crop = gather_crops(self._model, self.inimg, self.cs, self.ucs, self.ol, i, i + 1)[0]
        t = self.table[i]
        return crop, torch.IntTensor(t[2:6].tolist()), torch.IntTensor(t[6:8].tolist())
