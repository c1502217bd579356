"""CPU: the algebra behind DESIGN §9's "pixel-pair view" (not built yet) — a valid 3x3 convolution over an
NHWC tensor [H, W, C] equals a 3(y) x 2(x)-tap convolution over the same memory viewed as [H, W/2, 2C] with
N = 2*C_out outputs, where tap (ky, j) holds the 2x2 block matrix
    block(out pixel a, in pixel e) = W3x3[ky][2j + e - a]  if 0 <= 2j + e - a <= 2 else 0
(6 of 8 blocks non-zero).  This is what would let the C_out = 64 layers run as N = 128 GEMMs."""
import numpy as np
import torch
import torch.nn.functional as F


def pair_weights(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, 3, 3] -> [2*C_out, 2*C_in, 3, 2] acting on the pair view (channels = (pixel parity, c))."""
    co, ci = w.shape[:2]
    wp = torch.zeros(2 * co, 2 * ci, 3, 2, dtype=w.dtype)
    for j in range(2):
        for a in range(2):
            for e in range(2):
                kx = 2 * j + e - a
                if 0 <= kx <= 2:
                    wp[a * co:(a + 1) * co, e * ci:(e + 1) * ci, :, j] = w[:, :, :, kx]
    return wp


def test_pair_view_equals_3x3():
    g = torch.Generator().manual_seed(0)
    for (h, w, ci, co) in ((9, 12, 4, 3), (7, 20, 8, 8)):
        x = torch.rand(1, ci, h, w, generator=g, dtype=torch.float64)
        wt = torch.rand(co, ci, 3, 3, generator=g, dtype=torch.float64) - 0.5
        ref = F.conv2d(x, wt)                                             # [1, co, h-2, w-2]
        # pair view: [1, 2*ci, h, w/2], channel (e, c) = pixel 2p+e
        xp = x.view(1, ci, h, w // 2, 2).permute(0, 4, 1, 2, 3).reshape(1, 2 * ci, h, w // 2)
        yp = F.conv2d(xp, pair_weights(wt))                               # [1, 2*co, h-2, w/2-1]
        got = yp.view(1, 2, co, h - 2, w // 2 - 1).permute(0, 2, 3, 4, 1).reshape(1, co, h - 2, w - 2)
        assert float((got - ref).abs().max()) < 1e-12
    wp = pair_weights(torch.ones(1, 1, 3, 3))
    assert int((wp != 0).sum()) == 6 * 3      # 6 of 8 blocks per ky
