"""Scores the reference's batch path reports for every denoised image:

``pt_helpers.get_losses`` (/root/reference/src/nind_denoise/common/libs/pt_helpers.py:42-50) returns
``mse`` = ``F.mse_loss``, ``ssim`` = ``1 - piqa.SSIM()`` and ``msssim`` = ``1 - piqa.MS_SSIM()``
(common/libs/pt_losses.py:6-18) of the clean / denoised pair, and ``denoise_dir.py:99-129`` averages them and
records them under ``test_*`` keys in ``trainres.json`` / ``testres.json`` through ``JSONSaver``
(common/libs/json_saver.py:9-56).

``piqa`` (pinned ~=1.3.2 in the reference's pyproject.toml:35) is not a dependency here, so SSIM / MS-SSIM are
restated from their published definitions with that package's defaults: Gaussian window 11, sigma 1.5,
K1 = 0.01, K2 = 0.03, value range 1, "valid" (un-padded) separable filtering, per-channel maps averaged over
channels and pixels; MS-SSIM: five scales with weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), 2x2 average
pooling (ceil mode) between scales, contrast-structure terms clamped at 0 (Wang et al. 2003/2004).  They run on
whatever device the tensors live on (plain torch ops; scoring is not part of the timed hot path).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _gaussian(window: int = 11, sigma: float = 1.5, device=None) -> torch.Tensor:
    x = torch.arange(window, dtype=torch.float32, device=device) - (window - 1) / 2
    k = torch.exp(-(x ** 2) / (2 * sigma ** 2))
    return k / k.sum()


def _blur(x: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """Separable 'valid' Gaussian filtering of every channel of [B,C,H,W]."""
    c = x.shape[1]
    kh = k.view(1, 1, -1, 1).expand(c, 1, -1, 1)
    kw = k.view(1, 1, 1, -1).expand(c, 1, 1, -1)
    return F.conv2d(F.conv2d(x, kh, groups=c), kw, groups=c)


def _ssim_maps(x: torch.Tensor, y: torch.Tensor, k: torch.Tensor, value_range: float = 1.0, k1: float = 0.01,
               k2: float = 0.03):
    c1, c2 = (k1 * value_range) ** 2, (k2 * value_range) ** 2
    mu_x, mu_y = _blur(x, k), _blur(y, k)
    mu_xx, mu_yy, mu_xy = mu_x * mu_x, mu_y * mu_y, mu_x * mu_y
    s_xx = _blur(x * x, k) - mu_xx
    s_yy = _blur(y * y, k) - mu_yy
    s_xy = _blur(x * y, k) - mu_xy
    cs = (2 * s_xy + c2) / (s_xx + s_yy + c2)
    ss = (2 * mu_xy + c1) / (mu_xx + mu_yy + c1) * cs
    return ss, cs


def ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """SSIM of two [B,3,H,W] images in [0,1] -> [B] (mean over channels and pixels)."""
    ss, _ = _ssim_maps(x, y, _gaussian(device=x.device))
    return ss.flatten(1).mean(dim=1)


def ms_ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Multi-scale SSIM of two [B,3,H,W] images in [0,1] -> [B]; H, W >= 162 (pt_losses.py:20-29)."""
    k = _gaussian(device=x.device)
    w = torch.tensor(MS_WEIGHTS, dtype=torch.float32, device=x.device)
    terms = []
    for i in range(len(MS_WEIGHTS)):
        if i > 0:
            x = F.avg_pool2d(x, 2, ceil_mode=True)
            y = F.avg_pool2d(y, 2, ceil_mode=True)
        if min(x.shape[-2:]) < k.numel():
            raise RuntimeError(f"ms_ssim: image too small for {len(MS_WEIGHTS)} scales (needs >= 162 pixels per side)")
        ss, cs = _ssim_maps(x, y, k)
        t = ss if i + 1 == len(MS_WEIGHTS) else cs
        terms.append(torch.relu(t.flatten(2).mean(dim=2)))  # [B, C]
    m = torch.stack(terms, dim=-1) ** w  # [B, C, scales]
    return m.prod(dim=-1).mean(dim=-1)


def get_losses(clean: torch.Tensor, denoised: torch.Tensor) -> Dict[str, float]:
    """``pt_helpers.get_losses`` on two [3,H,W] tensors (the reference reads both from their files; values are
    whatever the files hold, no extra clamp): mse, ssim loss (1 - SSIM), msssim loss (1 - MS-SSIM)."""
    a, b = clean.unsqueeze(0).float(), denoised.unsqueeze(0).float()
    if a.shape != b.shape:
        raise ValueError(f"get_losses: shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    res = {"mse": float(F.mse_loss(a, b))}
    if min(a.shape[-2:]) >= 11:
        res["ssim"] = float(1 - ssim(a, b)[0])
    if min(a.shape[-2:]) >= 162:
        res["msssim"] = float(1 - ms_ssim(a, b)[0])
    return res


def avg_listofdicts(dicts):
    """Mean of every key over a list of dicts (what utilities.avg_listofdicts is meant to return; the fork's
    version at common/libs/utilities.py:61-69 forgets its return statement)."""
    dicts = [d for d in dicts if d]
    if not dicts:
        return {}
    return {k: sum(d[k] for d in dicts if k in d) / max(1, sum(1 for d in dicts if k in d)) for k in dicts[0]}


def add_test_results(json_path: str, epoch: Optional[int], res: Dict[str, float], key_prefix: str = "test_") -> dict:
    """``JSONSaver(json_path, step_type='epoch').add_res(step=epoch, res=res, key_prefix='test_')``
    (json_saver.py:9-52, denoise_dir.py:112-124): results keyed by epoch plus ``best_val`` / ``best_epoch`` (lower
    is better), merged into an existing file.  Same layout as the testres.json files the reference ships
    (src/nind_denoise/models/2021-05-31T22_11_nn_train/testres.json)."""
    data = {"best_val": {}}
    if os.path.isfile(json_path):
        with open(json_path) as fp:
            data = json.load(fp)
    data.setdefault("best_val", {})
    data.setdefault("best_epoch", {})
    step = str(epoch)
    data.setdefault(step, {})
    for k, v in res.items():
        key, v = key_prefix + k, float(v)
        data[step][key] = v
        if key not in data["best_val"] or data["best_val"][key] > v:
            data["best_val"][key] = v
            data["best_epoch"][key] = epoch
    with open(json_path, "w") as fp:
        json.dump(data, fp, indent=2)
    return data
