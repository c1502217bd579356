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


def plan_steps(model, width: int, height: int, cs: int, ucs: int, ol: int, crop_begin: int, crop_end: int,
               batch: int) -> List[Tuple[int, int]]:
    """The forwards a crop range is split into when copies / communication overlap the compute (``nind_plan_steps``:
    the first and last step end / start at a grid-row boundary, the rest are balanced forwards of <= batch crops)."""
    n = C.c_int()
    bounds = (C.c_int * 1026)()
    _capi.check(_capi.lib().nind_plan_steps(model.native_handle(), width, height, cs, ucs, ol, crop_begin, crop_end, batch,
                                            bounds, 1026, C.byref(n)))
    return [(bounds[i], bounds[i + 1]) for i in range(n.value)]


def tiled_step(model, img: torch.Tensor, out_img: torch.Tensor, cs: int, ucs: int, ol: int, crop_begin: int,
               crop_end: int, step_begin: int, step_end: int) -> Tuple[int, int]:
    """``nind_tiled_denoise_step``: one forward over crops [step_begin, step_end) of the range; the band rows that
    become final, [r0, r1), are stitched into ``out_img`` (full [3,H,W] layout, same device).  Returns (r0, r1)."""
    _, H, W = img.shape
    r0, r1 = C.c_int(), C.c_int()
    with torch.cuda.device(img.device):
        stream = torch.cuda.current_stream().cuda_stream
        _capi.check(_capi.lib().nind_tiled_denoise_step(model.native_handle(), img.data_ptr(), out_img.data_ptr(), H, W, cs,
                                                        ucs, ol, crop_begin, crop_end, step_begin, step_end, C.byref(r0),
                                                        C.byref(r1), C.c_void_p(stream)))
    return r0.value, r1.value


def add_rows(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst += src for two [3, rows, W] fp32 CUDA views whose planes are contiguous (``nind_add_rows``)."""
    assert dst.shape == src.shape and dst.dim() == 3 and dst.is_cuda and src.is_cuda
    if dst.numel() == 0:
        return
    assert dst[0].is_contiguous() and src[0].is_contiguous()
    count = dst.shape[1] * dst.shape[2]
    with torch.cuda.device(dst.device):
        _capi.check(_capi.lib().nind_add_rows(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0), dst.shape[0], count,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))


def copy_planes(dst: torch.Tensor, src: torch.Tensor, stream=None) -> None:
    """dst[...] = src[...] for two [3, rows, W] fp32 CUDA views with contiguous planes: ONE asynchronous 2-D copy on
    the source device's current stream (``nind_copy_planes``).  ``dst`` may be peer memory — rank 0's output image
    opened with ``nind_peer_open`` —: the copy is then a peer-to-peer DMA over NVLink."""
    assert dst.shape == src.shape and dst.dim() == 3 and dst.is_cuda and src.is_cuda
    if dst.numel() == 0:
        return
    assert dst[0].is_contiguous() and src[0].is_contiguous()
    count = dst.shape[1] * dst.shape[2]
    with torch.cuda.device(src.device):
        st = (stream or torch.cuda.current_stream()).cuda_stream
        _capi.check(_capi.lib().nind_copy_planes(dst.data_ptr(), dst.stride(0), src.data_ptr(), src.stride(0), dst.shape[0],
                                                 count, C.c_void_p(st)))


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


class _DevicePtr:
    """Zero-copy torch view of library-owned device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerGather:
    """Rank ``dst``'s output image(s) and seam scratch, mapped into every rank of the node through CUDA IPC, so that
    the gather of the stitched output is a set of peer-to-peer DMA copies over NVLink that each rank issues as
    soon as rows are final — on its copy engines, overlapped with the forwards still running, no SM taken from
    the persistent conv kernels (an NCCL send/recv kernel that waits for its peer would hold SMs the statically
    scheduled conv CTAs need).  NCCL is only used for the closing synchronisation.

    Two output images alternate between calls (a rank may start writing image k+1 while ``dst`` still finishes
    image k).  ``seam[r]`` receives rank r's partial sums for rows later ranks own (the grid row two crop ranges
    share, or the ``ol`` seam rows); ``dst`` adds them in rank = raster order.  Collective constructor."""

    def __init__(self, height: int, width: int, max_seam_rows: int, device, group=None, dst: int = 0):
        import torch.distributed as dist

        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.shape = (3, height, width)
        self.max_seam_rows = max(1, max_seam_rows)
        n_out = 2 * 3 * height * width
        n_seam = self.world * 3 * self.max_seam_rows * width
        lib = _capi.lib()
        ptr = C.c_void_p()
        payload = [None]
        self._owner = self.rank == dst
        with torch.cuda.device(device):
            if self._owner:
                handle = C.create_string_buffer(64)
                _capi.check(lib.nind_peer_alloc((n_out + n_seam) * 4, C.byref(ptr), handle))
                payload[0] = handle.raw
            src = dist.get_global_rank(group, dst) if group is not None else dst
            dist.broadcast_object_list(payload, src=src, group=group)
            if not self._owner:  # map rank dst's buffer on THIS rank's device (lazy peer access over NVLink)
                _capi.check(lib.nind_peer_open(payload[0], C.byref(ptr)))
        self._ptr = ptr.value
        base = torch.as_tensor(_DevicePtr(self._ptr, (n_out + n_seam,)), device=device)
        self.outs = base[:n_out].view(2, 3, height, width)
        self.seam = base[n_out:].view(self.world, 3, self.max_seam_rows, width)
        self.calls = 0
        dist.barrier(group)

    def close(self):
        if getattr(self, "_ptr", None):
            lib = _capi.lib()
            (lib.nind_peer_free if self._owner else lib.nind_peer_close)(C.c_void_p(self._ptr))
            self._ptr = None

    def out(self) -> torch.Tensor:
        return self.outs[self.calls & 1]


_peer_cache = {}


def _peer_gather(height, width, max_seam_rows, device, group, dst) -> PeerGather:
    key = (height, width, str(device), id(group), dst)
    pg = _peer_cache.get(key)
    if pg is None or pg.max_seam_rows < max_seam_rows:
        pg = PeerGather(height, width, max_seam_rows, device, group, dst)
        _peer_cache[key] = pg
    return pg


def _peer_seams(width, max_rows, device, group) -> "PeerSeams":
    key = ("seams", width, str(device), id(group))
    ps = _peer_cache.get(key)
    if ps is None or ps.max_rows < max_rows:
        ps = PeerSeams(width, max_rows, device, group)
        _peer_cache[key] = ps
    return ps


def _denoise_tiled_distributed_peer(img, model, cs, ucs, ol, batch, group, dst):
    """mode="peer" of ``denoise_tiled_distributed`` (see there)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _, H, W = img.shape
    n = n_crops(W, H, cs, ucs, ol)
    ranges = shard_ranges(n, world)
    cb, ce = ranges[rank]
    extents = band_extents(W, H, cs, ucs, ol, ranges)
    own = owned_rows(extents, H)
    max_seam = max([max(0, e[1] - o[1]) for e, o in zip(extents, own)] + [1])
    pg = _peer_gather(H, W, max_seam, img.device, group, dst)
    out = pg.out()
    dev = img.device
    key = ("full", H, W, str(dev))
    full = _peer_cache.get(key)
    if full is None:
        full = _peer_cache[key] = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    if ("side", str(dev)) not in _peer_cache:
        _peer_cache[("side", str(dev))] = torch.cuda.Stream(device=dev)
        _peer_cache[("token", str(dev))] = torch.zeros(1, device=dev)
    side = _peer_cache[("side", str(dev))]
    main = torch.cuda.current_stream(dev)
    if ce > cb:
        if batch is None:
            batch = default_batch(ce - cb, cs, _nx(W, ucs, ol))
        y0, y1 = extents[rank]
        o0, o1 = own[rank]
        # at least two balanced forwards, so that half of the copies overlap compute; small forwards are
        # inefficient (per-launch prologues, wave quantisation), so no finer than the batch asks for
        ncr = ce - cb
        k = max(2 if ncr >= 32 else 1, -(-ncr // batch))
        for i in range(k):
            a, b = cb + ncr * i // k, cb + ncr * (i + 1) // k
            r0, r1 = tiled_step(model, img, full, cs, ucs, ol, cb, ce, a, b)
            if r1 <= r0:
                continue
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            a0, a1 = max(r0, o0), min(r1, o1)          # rows this rank owns: straight into the output image
            if a1 > a0:
                copy_planes(out[:, a0:a1], full[:, a0:a1], side)
            s0, s1 = max(r0, o1), min(r1, y1)          # rows later ranks own: partial sums -> seam scratch
            if s1 > s0:
                copy_planes(pg.seam[rank, :, s0 - o1:s1 - o1], full[:, s0:s1], side)
        main.wait_stream(side)
    # closing synchronisation: every rank's copies are ordered before its contribution to this all-reduce
    token = _peer_cache[("token", str(dev))]
    dist.all_reduce(token, group=group)
    pg.calls += 1
    if rank != dst:
        return None
    for r in range(world):          # partial sums for rows later ranks own, in rank (= raster) order
        rows = extents[r][1] - own[r][1]
        if extents[r][1] > extents[r][0] and rows > 0:
            add_rows(out[:, own[r][1]:extents[r][1]], pg.seam[r, :, :rows])
    return out


def denoise_tiled_distributed(img: torch.Tensor, model, cs: int, ucs: int, ol: int = DEFAULT_OVERLAP,
                              batch: Optional[int] = None, group=None, dst: int = 0,
                              band_fn: Optional[Callable] = None, mode: str = "rows") -> Optional[torch.Tensor]:
    """Every rank holds the same ``img`` (read-only) and a replica of ``model``; rank ``dst`` returns the
    stitched image, the others return None.

    ``mode="peer"``: every rank copies the rows it owns straight into ``dst``'s output image, which is mapped
    into every rank through CUDA IPC (``PeerGather``), step by step as they become final — peer DMA over NVLink
    overlapped with the remaining forwards; partial sums for rows later ranks own go to a scratch area that
    ``dst`` adds at the end; the only collective is the closing synchronisation.  The returned image is one of two
    buffers that alternate between calls.
    ``mode="rows"`` (default): neighbours first exchange the seam rows they share (``exchange_seams``),
    then every rank sends only the rows it owns and ``dst`` receives them straight into the output image —
    each output byte crosses NVLink once and nothing is summed on ``dst`` (NCCL send / recv).
    ``mode="bands"``: whole overlapping bands are gathered and summed on ``dst`` (kept for comparison).

    ``band_fn(img, crop_begin, crop_end) -> (band, y0, y1)`` replaces the GPU band computation in CPU
    (gloo) tests of the sharding/gather logic."""
    import torch.distributed as dist

    if mode == "peer":
        if band_fn is not None:
            raise ValueError("mode='peer' runs on GPUs only (CUDA IPC)")
        return _denoise_tiled_distributed_peer(img, model, cs, ucs, ol, batch, group, dst)
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


def owned_rows_up(extents: Sequence[Tuple[int, int]], height: int) -> List[Tuple[int, int]]:
    """Row ownership of the host-buffer entry: a non-empty rank owns the rows from the END of the previous
    non-empty band up to the end of its own band (the last one up to ``height``).  The grid row two crop ranges
    share therefore belongs to the EARLIER rank — the later rank computes its part of that row first (raster
    order) and can hand it over while both still run, whereas the earlier rank finishes its part last."""
    own = [(0, 0)] * len(extents)
    live = [r for r, (y0, y1) in enumerate(extents) if y1 > y0]
    prev_end = 0
    for i, r in enumerate(live):
        end = max(prev_end, extents[r][1] if i + 1 < len(live) else height)
        own[r] = (prev_end, end)
        prev_end = end
    return own


def seam_plan(extents, own, rank: int):
    """(sends, recvs) of ``rank`` for an ownership table: sends = [(owner, a, b)] rows [a, b) of this rank's band that
    another rank owns; recvs = [(sender, a, b)] rows of this rank's own range that rank ``sender``'s band also
    touches — in increasing rank order, which is the order the partial sums are added in."""
    y0, y1 = extents[rank]
    o0, o1 = own[rank]
    sends, recvs = [], []
    for r in range(len(extents)):
        if r == rank:
            continue
        a, b = max(y0, own[r][0]), min(y1, own[r][1])
        if y1 > y0 and b > a:
            sends.append((r, a, b))
        a, b = max(extents[r][0], o0), min(extents[r][1], o1)
        if extents[r][1] > extents[r][0] and b > a:
            recvs.append((r, a, b))
    return sends, recvs


def exchange_seams_up(band: Optional[torch.Tensor], extents, own, rank: int, group=None) -> None:
    """Blocking form of the seam hand-over for any ownership table (CPU / gloo tests; the GPU path uses
    ``PeerSeams``): one batched P2P group, partial sums added in rank order."""
    import torch.distributed as dist

    if band is None:
        return
    y0 = extents[rank][0]
    sends, recvs = seam_plan(extents, own, rank)
    ops, keep, bufs = [], [], []
    for r, a, b in sends:
        keep.append(band[:, a - y0:b - y0, :].contiguous())
        ops.append(dist.P2POp(dist.isend, keep[-1], r, group))
    for r, a, b in recvs:
        bufs.append(torch.empty((3, b - a, band.shape[2]), dtype=band.dtype, device=band.device))
        ops.append(dist.P2POp(dist.irecv, bufs[-1], r, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for (r, a, b), buf in zip(recvs, bufs):
        band[:, a - y0:b - y0, :] += buf


class PeerSeams:
    """Every rank's seam receive area ([2 images][world senders][3, max_rows, W] fp32), mapped into the ranks that
    send to it through CUDA IPC: a rank hands the rows a neighbour owns to that neighbour with ONE peer-to-peer DMA
    over NVLink as soon as its own crops are done with them (copy engines; no NCCL kernel that would take SMs from
    the persistent conv kernels and wait for its peer).  Completion is signalled through ``SharedHostImage``'s
    flags.  Collective constructor."""

    def __init__(self, width: int, max_rows: int, device, group=None):
        import torch.distributed as dist

        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.width, self.max_rows, self.device = width, max(1, max_rows), device
        self.slot = 3 * self.max_rows * width
        lib = _capi.lib()
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(device):
            _capi.check(lib.nind_peer_alloc(2 * self.world * self.slot * 4, C.byref(ptr), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self._handles = handles
        self._ptrs = {self.rank: ptr.value}
        dist.barrier(group)

    def _base(self, owner: int) -> int:
        if owner not in self._ptrs:  # map the owner's area on THIS rank's device (lazy peer access over NVLink)
            ptr = C.c_void_p()
            with torch.cuda.device(self.device):
                _capi.check(_capi.lib().nind_peer_open(self._handles[owner], C.byref(ptr)))
            self._ptrs[owner] = ptr.value
        return self._ptrs[owner]

    def view(self, owner: int, parity: int, sender: int, rows: int) -> torch.Tensor:
        """[3, rows, W] view of rank ``owner``'s slot for ``sender`` (image parity 0 | 1), on this rank's device."""
        assert rows <= self.max_rows
        off = ((parity & 1) * self.world + sender) * self.slot * 4
        t = torch.as_tensor(_DevicePtr(self._base(owner) + off, (3, self.max_rows, self.width)), device=self.device)
        return t[:, :rows]

    def close(self):
        lib = _capi.lib()
        for owner, ptr in list(getattr(self, "_ptrs", {}).items()):
            (lib.nind_peer_free if owner == self.rank else lib.nind_peer_close)(C.c_void_p(ptr))
        self._ptrs = {}


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


def bind_host_to_gpu(device_index: int) -> Optional[List[int]]:
    """Pin the calling process to the CPUs of the NUMA node GPU ``device_index`` hangs off (sysfs ``local_cpulist`` of
    its PCI function), so that the host buffers it allocates AFTERWARDS (first touch) and the thread that feeds the
    GPU sit next to its PCIe root port.  One process per GPU (torchrun) does not do this by itself.  Returns the CPU
    list, or None when the topology cannot be read (then nothing is changed)."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        cpus: List[int] = []
        for part in open(path).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.extend(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class SharedHostImage:
    """A [3,H,W] fp32 host image in ONE shared-memory segment mapped by every rank of ``group`` (ranks of a
    node), page-locked on each rank: the multi-GPU host entry lets every GPU copy the rows it owns straight
    into it over its own PCIe link instead of funnelling the whole image through rank ``src``'s.
    Collective constructor.  ``tensor`` is the CPU view (all ranks see the same bytes).

    The segment also carries one 64-bit arrival counter per rank (``arrive`` / ``wait_all``): the closing
    synchronisation of an image — "every rank's rows have landed" — is a store by each rank and a spin over
    ``world`` cache lines by the reader, instead of an NCCL barrier (a kernel launch and a stream sync per rank)."""

    _FLAG_STRIDE = 16  # int64s per rank: one 128-byte line each, no false sharing

    def __init__(self, shape, group=None, src: int = 0, pin: bool = True):
        import os

        import torch.distributed as dist

        self.shape = tuple(int(v) for v in shape)
        numel = math.prod(self.shape)
        self._fd = None
        self._group = group
        rank = dist.get_rank(group)
        self._rank, self._world = rank, dist.get_world_size(group)
        flag_bytes = self._world * self._FLAG_STRIDE * 8
        data_bytes = (numel * 4 + 4095) // 4096 * 4096
        path = [None]
        if rank == src:
            # anonymous memory file, reachable by the other ranks through /proc (no /dev/shm size limit)
            self._fd = os.memfd_create("nind_b200_out")
            os.ftruncate(self._fd, data_bytes + flag_bytes)
            path[0] = f"/proc/{os.getpid()}/fd/{self._fd}"
        dist.broadcast_object_list(path, src=dist.get_global_rank(group, src) if group is not None else src,
                                   group=group)
        whole = torch.from_file(path[0], shared=True, size=data_bytes + flag_bytes, dtype=torch.uint8)
        self._whole = whole
        self.tensor = whole[:numel * 4].view(torch.float32).view(self.shape)
        self._flags = whole[data_bytes:].view(torch.int64)
        if rank == src:
            self._flags.zero_()
        # First touch: every rank writes the slice of rows it will (roughly) own before anyone page-locks the segment,
        # so that those pages are allocated on ITS NUMA node (ownership is by contiguous row ranges in rank order).
        if len(self.shape) == 3:
            hh = self.shape[1]
            self.tensor[:, hh * rank // self._world:hh * (rank + 1) // self._world].zero_()
        dist.barrier(group)
        self._seq = 0
        self.pinned = False
        if pin and torch.cuda.is_available():
            _capi.check(_capi.lib().nind_host_register(self.tensor.data_ptr(), numel * 4))
            self.pinned = True
        dist.barrier(group)  # everyone has mapped it; the creator can drop its descriptor
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None

    def arrive(self) -> int:
        """This rank's rows of the current image are in the segment (call after the copies were synchronised)."""
        self._seq += 1
        self._flags[self._rank * self._FLAG_STRIDE] = self._seq
        return self._seq

    def seam_rows_sent(self, seq: int, rows: int) -> None:
        """``rows`` of the rows this rank hands over for image ``seq`` have landed in their owners' receive areas."""
        self._flags[self._rank * self._FLAG_STRIDE + 1] = (seq << 32) | rows

    def seam_rows_of(self, rank: int, seq: int) -> int:
        """How many of rank ``rank``'s hand-over rows of image ``seq`` have landed (all of them once it is past it)."""
        v = int(self._flags[rank * self._FLAG_STRIDE + 1])
        return (v & 0xFFFFFFFF) if (v >> 32) == seq else (1 << 31 if (v >> 32) > seq else 0)

    def wait_flag(self, rank: int, slot: int, seq: int, timeout: float = 60.0) -> None:
        """Spin until rank ``rank``'s counter ``slot`` (0 arrival, 1 seam rows sent) has reached ``seq``."""
        import time

        view = self._flags[rank * self._FLAG_STRIDE + slot]
        t0 = time.perf_counter()
        while int(view) < seq:
            if time.perf_counter() - t0 > timeout:
                raise RuntimeError(f"SharedHostImage: rank {rank} did not reach image {seq} within the time-out")

    def wait_all(self, timeout: float = 60.0) -> None:
        """Spin until every rank has arrived at this rank's current sequence number."""
        import time

        view = self._flags[::self._FLAG_STRIDE][:self._world]
        t0 = time.perf_counter()
        while int(view.min()) < self._seq:
            if time.perf_counter() - t0 > timeout:
                raise RuntimeError("SharedHostImage.wait_all: a rank did not arrive within the time-out")

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
                                   band_fn: Optional[Callable] = None,
                                   phases: Optional[list] = None) -> Optional[torch.Tensor]:
    """Multi-GPU host-buffer entry: every rank holds the same CPU image (pinned for full PCIe speed),
    uploads only the rows its crop range reads and stitches its band on its GPU.

    * ``out`` a ``SharedHostImage``: every rank copies the rows it owns (``owned_rows_up``) straight into the
      shared host image — N PCIe links in parallel, no gather.  The rows of its band an earlier rank owns (its
      part of the grid row the two crop ranges share) are final after its first step(s) and go to that rank's
      GPU by peer DMA (``PeerSeams``) while the remaining crops run; the owner adds them to its own partial
      sums after its last step — the step that produces those rows — and downloads them with that step's rows.
      ``phases`` (a list) receives (name, host time) marks of the call's own blocking points.
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
        own = owned_rows_up(extents, H)
        y0, y1 = extents[rank]
        o0, o1 = own[rank]
        sends, recvs = seam_plan(extents, own, rank)
        # own rows that later ranks' bands also touch are only final after their partial sums were added
        hi = max(o0, min([o1] + [a for _, a, _ in recvs]))
        seq = out._seq + 1  # the image this call produces
        import time as _time

        def mark(name):
            if phases is not None:
                phases.append((name, _time.perf_counter()))

        mark("start")
        if band_fn is not None:  # CPU (gloo) emulation: blocking hand-over at the end
            band = None
            if ce > cb:
                band, by0, by1 = band_fn(img_host, cb, ce)
                assert (by0, by1) == (y0, y1)
            exchange_seams_up(band, extents, own, rank, group)
            if o1 > o0:
                out.tensor[:, o0:o1, :].copy_(band[:, o0 - y0:o1 - y0, :])
        else:
            dev = model._device if getattr(model, "_handle", None) else next(model.parameters()).device
            max_rows = max([b - a for r in range(world) for _, a, b in seam_plan(extents, own, r)[0]] + [1])
            ps = _peer_seams(W, max_rows, dev, group)  # collective on first use: every rank, also one without crops
        if band_fn is not None:
            pass
        elif ce > cb:
            if batch is None:
                batch = default_batch(ce - cb, cs, _nx(W, ucs, ol))
            if ("side", str(dev)) not in _peer_cache:
                _peer_cache[("side", str(dev))] = torch.cuda.Stream(device=dev)
                _peer_cache[("token", str(dev))] = torch.zeros(1, device=dev)
            side = _peer_cache[("side", str(dev))]
            # H2D | forward | stitch | D2H of rows [o0, hi) pipelined inside the library
            full = host_range(model, img_host, out.tensor, cs, ucs, ol, batch, cb, ce, o0, hi)
            mark("enqueued")
            lib, h = _capi.lib(), model.native_handle()
            # Hand-over: the rows of this band that earlier ranks own — its part of the grid row it shares with the
            # previous rank — become final step by step (all but the last `ol` rows after the first step); each
            # chunk goes to its owner's receive area by peer DMA on a side stream that only waits for that step.
            chunks = []  # (event, rows handed over so far)
            if sends:
                for r, _, _ in sends:  # the owner must be done with image seq-2, which used the same slot
                    out.wait_flag(r, 0, seq - 2)
                nst = C.c_int()
                rd = (C.c_int * 1024)()
                _capi.check(lib.nind_host_rows_done(h, rd, 1024, C.byref(nst)))
                prev, total = y0, 0
                for k in range(nst.value):
                    todo = [(r, a, max(a, prev), min(b, rd[k])) for r, a, b in sends if min(b, rd[k]) > max(a, prev)]
                    if todo:
                        with torch.cuda.device(dev):
                            _capi.check(lib.nind_host_join_rows(h, rd[k], C.c_void_p(side.cuda_stream)))
                        for r, a, c0, c1 in todo:
                            copy_planes(ps.view(r, seq, rank, c1 - a)[:, c0 - a:], full[:, c0:c1], side)
                            total += c1 - c0
                        ev = torch.cuda.Event()
                        ev.record(side)
                        chunks.append((ev, total))
                    prev = max(prev, rd[k])
            # Receive side, in rank order: add what the later ranks hand over as it arrives (their counters say how
            # many rows have landed) and download the own rows below `hi` as far as they are final.
            state = []  # per sender: [rank, a, b, rows before this range in the sender's hand-over order, added]
            for r, a, b in recvs:
                before = sum(bb - aa for _, aa, bb in seam_plan(extents, own, r)[0] if aa < a)
                state.append([r, a, b, before, 0])
            landed = hi   # own rows [o0, landed) are on their way to the shared host image
            cur, spins, t_spin = 0, 0, _time.perf_counter()
            while chunks or cur < len(state) or landed < o1:
                spins += 1
                if (spins & 0xFFF) == 0 and _time.perf_counter() - t_spin > 60.0:
                    raise RuntimeError("denoise_tiled_distributed_host: a neighbour's rows did not arrive within 60 s")
                if chunks and chunks[0][0].query():
                    out.seam_rows_sent(seq, chunks.pop(0)[1])
                    if not chunks:
                        mark("handed_over")
                if cur < len(state):
                    r, a, b, before, added = state[cur]
                    have = min(b - a, max(0, out.seam_rows_of(r, seq) - before))
                    if have > added:
                        add_rows(full[:, a + added:a + have], ps.view(rank, seq, r, have)[:, added:])
                        state[cur][4] = have
                        if have == b - a:
                            cur += 1
                            if cur == len(state):
                                mark("received")
                final = min([o1] + [st[1] + st[4] for st in state[cur:]])
                if final > landed:
                    for c in range(3):
                        out.tensor[c, landed:final].copy_(full[c, landed:final], non_blocking=True)
                    landed = final
            if not sends:
                mark("handed_over")
            _capi.check(_capi.lib().nind_host_sync(model.native_handle()))
            torch.cuda.current_stream(dev).synchronize()
            mark("rows_landed")
        # every rank's rows have landed in the shared image: arrival counters in the segment itself (no NCCL
        # barrier); only the rank that returns the image waits
        out.arrive()
        if rank == dst:
            out.wait_all()
        mark("all_arrived")
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
