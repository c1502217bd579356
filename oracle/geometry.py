"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement, in numpy, of the crop grid / mirror-padded crop gather / trim / seam-halving /
overlap-add geometry of the reference tiler.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline leg may import this module.

Reference (paths relative to /root/reference):
  * grid            src/nind_denoise/denoise_image.py:88-107   (OneImageDS.__init__)
  * crop gather     src/nind_denoise/denoise_image.py:129-174  (OneImageDS.__getitem__, tiled branch)
  * trim            src/nind_denoise/denoise_image.py:250-254
  * seam halving    src/nind_denoise/denoise_image.py:204-213  (make_seamless_edges)
  * overlap-add     src/nind_denoise/denoise_image.py:267

Pinned by tests/golden/geometry_*.npz, which were produced by running the reference's own
OneImageDS.__getitem__ (oracle/make_golden.py) — the reference ships no golden vectors itself.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class CropGrid:
    """Integer description of the crop grid (denoise_image.py:100-107)."""

    width: int
    height: int
    cs: int
    ucs: int
    ol: int
    pad: int
    stride: int
    nx: int  # crops per line  = iperhl + 1
    ny: int  # lines of crops  = ipervl + 1

    @property
    def size(self) -> int:
        return self.nx * self.ny


def crop_grid(width: int, height: int, cs: int, ucs: int, ol: int) -> CropGrid:
    """denoise_image.py:100-104: iperhl = ceil((W-ucs)/(ucs-ol)), pad = int((cs-ucs)/2), ..."""
    stride = ucs - ol
    iperhl = math.ceil((width - ucs) / stride)
    ipervl = math.ceil((height - ucs) / stride)
    pad = int((cs - ucs) / 2)
    return CropGrid(width, height, cs, ucs, ol, pad, stride, iperhl + 1, ipervl + 1)


def crop_entry(g: CropGrid, i: int) -> dict:
    """Indices of crop ``i`` (denoise_image.py:130-137,139-143,172-173).

    Returns x0,y0 (top-left of the cs x cs window in image coordinates, may be negative), the
    four pad amounts, ``usefuldim`` = (x_lo, y_lo, x_hi, y_hi) inside the crop and ``usefulstart``
    = (x, y) of the useful area in the image.
    """
    yi = int(math.ceil((i + 1) / g.nx - 1))
    xi = i - yi * g.nx
    x0 = g.ucs * xi - g.ol * xi - g.pad
    y0 = g.ucs * yi - g.ol * yi - g.pad
    x1, y1 = x0 + g.cs, y0 + g.cs
    x0pad, x1pad = -min(0, x0), max(0, x1 - g.width)
    y0pad, y1pad = -min(0, y0), max(0, y1 - g.height)
    usefuldim = (g.pad, g.pad, g.cs - max(g.pad, x1pad), g.cs - max(g.pad, y1pad))
    usefulstart = (x0 + g.pad, y0 + g.pad)
    return dict(xi=xi, yi=yi, x0=x0, y0=y0, x0pad=x0pad, x1pad=x1pad, y0pad=y0pad, y1pad=y1pad,
                usefuldim=usefuldim, usefulstart=usefulstart)


def crop_table(g: CropGrid) -> np.ndarray:
    """int32 [size, 8]: x0, y0, ud_xlo, ud_ylo, ud_xhi, ud_yhi, start_x, start_y per crop."""
    t = np.zeros((g.size, 8), dtype=np.int32)
    for i in range(g.size):
        e = crop_entry(g, i)
        t[i] = (e["x0"], e["y0"], *e["usefuldim"], *e["usefulstart"])
    return t


def _sym(t: np.ndarray, n: int) -> np.ndarray:
    """Edge-inclusive mirror index: the reference fills the out-of-image parts of a crop with
    np.flip of the adjacent image block (denoise_image.py:151-170), i.e. index -1 -> 0,
    -2 -> 1, n -> n-1, n+1 -> n-2."""
    t = np.where(t < 0, -t - 1, t)
    return np.where(t >= n, 2 * n - 1 - t, t)


def gather_crop(img: np.ndarray, g: CropGrid, i: int) -> np.ndarray:
    """Crop ``i`` of a CHW float32 image with mirror padding (denoise_image.py:138-170)."""
    e = crop_entry(g, i)
    ys = _sym(np.arange(e["y0"], e["y0"] + g.cs), g.height)
    xs = _sym(np.arange(e["x0"], e["x0"] + g.cs), g.width)
    return np.ascontiguousarray(img[:, ys[:, None], xs[None, :]])


def seam_weights(g: CropGrid, i: int) -> np.ndarray:
    """Per-pixel weight the reference applies to the trimmed crop ``i`` before adding it
    (make_seamless_edges, denoise_image.py:204-213): halve the first/last ``ol`` columns/rows
    when a neighbour exists on that side; corners end up at 1/4."""
    e = crop_entry(g, i)
    xlo, ylo, xhi, yhi = e["usefuldim"]
    h, w = yhi - ylo, xhi - xlo
    ax, ay = e["usefulstart"]
    wt = np.ones((h, w), dtype=np.float32)
    if ax != 0:
        wt[:, 0:g.ol] /= 2
    if ay != 0:
        wt[0:g.ol, :] /= 2
    if ax + g.ucs < g.width and g.ol:
        wt[:, -g.ol:] /= 2
    if ay + g.ucs < g.height and g.ol:
        wt[-g.ol:, :] /= 2
    return wt


def stitch(crops_out, g: CropGrid, crop_range=None) -> np.ndarray:
    """Trim + seam-halve + overlap-add in raster crop order (denoise_image.py:250-267).

    ``crops_out``: callable i -> [3, cs, cs] float32 network output of crop i, or an indexable.
    """
    out = np.zeros((3, g.height, g.width), dtype=np.float32)
    rng = range(g.size) if crop_range is None else range(*crop_range)
    for i in rng:
        e = crop_entry(g, i)
        xlo, ylo, xhi, yhi = e["usefuldim"]
        ax, ay = e["usefulstart"]
        full = crops_out(i) if callable(crops_out) else crops_out[i]
        t = np.array(full[:, ylo:yhi, xlo:xhi], dtype=np.float32, copy=True)
        # the reference halves in place, side by side (left, top, right, bottom)
        if ax != 0:
            t[:, :, 0:g.ol] = t[:, :, 0:g.ol] / 2
        if ay != 0:
            t[:, 0:g.ol, :] = t[:, 0:g.ol, :] / 2
        if ax + g.ucs < g.width and g.ol:
            t[:, :, -g.ol:] = t[:, :, -g.ol:] / 2
        if ay + g.ucs < g.height and g.ol:
            t[:, -g.ol:, :] = t[:, -g.ol:, :] / 2
        out[:, ay:ay + t.shape[1], ax:ax + t.shape[2]] += t
    return out


def denoise_tiled(img: np.ndarray, model_fn, cs: int, ucs: int, ol: int, crop_range=None) -> np.ndarray:
    """The reference main loop (denoise_image.py:231-267) with ``model_fn`` standing for the
    network: CHW float32 crop [3,cs,cs] -> [3,cs,cs]."""
    g = crop_grid(img.shape[2], img.shape[1], cs, ucs, ol)
    return stitch(lambda i: model_fn(gather_crop(img, g, i)), g, crop_range)


def whole_image_input(img: np.ndarray, pad: int) -> np.ndarray:
    """The whole-image branch of OneImageDS.__getitem__ (denoise_image.py:110-126): image centred in a zero
    canvas, four sides mirrored (edge pixel included), corners left zero.  The reference allocates the canvas
    as (3, W+2p, H+2p) (:113), which only works for square images; this is the (3, H+2p, W+2p) layout."""
    _, H, W = img.shape
    ret = np.zeros((3, H + 2 * pad, W + 2 * pad), dtype=np.float32)
    ret[:, pad:H + pad, pad:W + pad] = img
    if pad:
        ret[:, pad:-pad, :pad] = np.flip(img[:, :, :pad], axis=2)
        ret[:, pad:-pad, W + pad:] = np.flip(img[:, :, W - pad:], axis=2)
        ret[:, :pad, pad:-pad] = np.flip(img[:, :pad, :], axis=1)
        ret[:, H + pad:, pad:-pad] = np.flip(img[:, H - pad:, :], axis=1)
    return ret
