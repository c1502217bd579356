"""ORACLE (test infrastructure only — never imported by the product path).

fp32 CPU restatement of the two NIND denoiser networks as pure functions of a ``state_dict``
(torch CPU ops; the reference's arithmetic is torch's own conv/pool/prelu, pin torch~=2.8 in
/root/reference/pyproject.toml:46, container has torch 2.11).

Reference:
  * UtNet   src/nind_denoise/networks/UtNet.py:14-109
  * UNet    src/nind_denoise/networks/ThirdPartyNets.py:62-169

``init_state_dict`` re-creates the reference modules' default initialisation (same layer creation
order, hence the same RNG stream under ``torch.manual_seed``) so that tests on the GPU box — where
/root/reference does not exist — use bit-identical weights to the ones the golden vectors in
tests/golden/ were generated with (oracle/make_golden.py checks the state_dicts are equal).
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

# ----------------------------------------------------------------------------- layer tables
# (state_dict prefix, kind, c_in multiplier, c_out multiplier) in the reference's creation order.
# kinds: c3 = Conv2d k3, t3 = ConvTranspose2d k3, t2 = ConvTranspose2d k2 s2, c1 = Conv2d k1 -> 3 ch


def utnet_layers(funit: int = 64):
    f = funit
    L = []
    enc = [(3, f), (f, 2 * f), (2 * f, 4 * f), (4 * f, 8 * f)]
    for lvl, (ci, co) in enumerate(enc, start=1):
        L.append((f"convs{lvl}.0", "c3", ci, co))
        L.append((f"convs{lvl}.2", "c3", co, co))
    L.append(("bottom.0", "c3", 8 * f, 16 * f))
    L.append(("bottom.2", "t3", 16 * f, 16 * f))
    width = 16 * f
    for lvl in range(1, 5):
        half = width // 2
        L.append((f"up{lvl}", "t2", width, half))
        L.append((f"tconvs{lvl}.0", "t3", width, half))
        L.append((f"tconvs{lvl}.2", "t3", half, half))
        width = half
    L.append(("tconvs4.4", "c1", f, 3))
    return L


UTNET_ACT_SLOTS = ([f"convs{i}.{j}" for i in range(1, 5) for j in (1, 3)] + ["bottom.1", "bottom.3"] +
                   [f"tconvs{i}.{j}" for i in range(1, 5) for j in (1, 3)])


def unet_layers():
    """(prefix, kind, cin, cout); 'bn' rows are BatchNorm2d(cout).  ThirdPartyNets.py:139-150."""
    L = []

    def dconv(prefix, ci, co):
        L.append((prefix + ".0", "c3p", ci, co))
        L.append((prefix + ".1", "bn", co, co))
        L.append((prefix + ".3", "c3p", co, co))
        L.append((prefix + ".4", "bn", co, co))

    dconv("inc.conv.conv", 3, 64)
    for i, (ci, co) in enumerate([(64, 128), (128, 256), (256, 512), (512, 512)], start=1):
        dconv(f"down{i}.mpconv.1.conv", ci, co)
    for i, (ci, co) in enumerate([(1024, 256), (512, 128), (256, 64), (128, 64)], start=1):
        L.append((f"up{i}.up", "t2", ci // 2, ci // 2))
        dconv(f"up{i}.conv.conv", ci, co)
    L.append(("outc.conv", "c1", 64, 3))
    return L


def _make(kind, ci, co):
    if kind == "c3":
        return nn.Conv2d(ci, co, 3)
    if kind == "c3p":
        return nn.Conv2d(ci, co, 3, padding=1)
    if kind == "t3":
        return nn.ConvTranspose2d(ci, co, 3)
    if kind == "t2":
        return nn.ConvTranspose2d(ci, co, 2, stride=2)
    if kind == "c1":
        return nn.Conv2d(ci, co, 1)
    if kind == "bn":
        return nn.BatchNorm2d(co)
    raise ValueError(kind)


def init_state_dict(network: str = "UtNet", seed: int = 0, funit: int = 64, activation: str = "PReLU"):
    """Default-initialised weights, identical to ``torch.manual_seed(seed); <reference class>()``."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    if network == "UtNet":
        mods = OrderedDict((name, _make(kind, ci, co)) for name, kind, ci, co in utnet_layers(funit))
        # state_dict order follows attribute registration order of the reference module
        order = []
        for lvl in range(1, 5):
            order += [f"convs{lvl}.0", f"convs{lvl}.1", f"convs{lvl}.2", f"convs{lvl}.3"]
        order += ["bottom.0", "bottom.1", "bottom.2", "bottom.3"]
        for lvl in range(1, 5):
            order += [f"up{lvl}", f"tconvs{lvl}.0", f"tconvs{lvl}.1", f"tconvs{lvl}.2", f"tconvs{lvl}.3"]
        order += ["tconvs4.4"]
        for name in order:
            if name in mods:
                sd[name + ".weight"] = mods[name].weight.detach().clone()
                sd[name + ".bias"] = mods[name].bias.detach().clone()
            elif activation == "PReLU":
                sd[name + ".weight"] = torch.full((1,), 0.25)
    elif network == "UNet":
        for name, kind, ci, co in unet_layers():
            m = _make(kind, ci, co)
            for k, v in m.state_dict().items():
                sd[f"{name}.{k}"] = v.detach().clone()
    else:
        raise ValueError(network)
    return sd


def randomize_bn_(sd, seed: int = 7):
    """Non-trivial BatchNorm statistics/affine (at init BN is the identity)."""
    g = torch.Generator().manual_seed(seed)
    for k in list(sd.keys()):
        if k.endswith("running_mean"):
            p = k[: -len("running_mean")]
            n = sd[k].numel()
            sd[p + "running_mean"] = torch.randn(n, generator=g) * 0.1
            sd[p + "running_var"] = torch.rand(n, generator=g) * 0.5 + 0.75
            sd[p + "weight"] = torch.rand(n, generator=g) * 0.5 + 0.75
            sd[p + "bias"] = torch.randn(n, generator=g) * 0.1
    return sd


# ----------------------------------------------------------------------------- forwards
def _act(x, sd, key, activation):
    if activation == "PReLU":
        return F.prelu(x, sd[key + ".weight"])
    if activation == "ELU":
        return F.elu(x)
    if activation == "Hardswish":
        return F.hardswish(x)
    raise ValueError(activation)


def utnet_forward(sd, x: torch.Tensor, activation: str = "PReLU") -> torch.Tensor:
    """UtNet.forward (UtNet.py:97-109) on [B,3,cs,cs] fp32, cs = 16a+56 (a >= 3)."""
    w = lambda k: sd[k + ".weight"]
    b = lambda k: sd[k + ".bias"]
    conv = lambda t, k: F.conv2d(t, w(k), b(k))
    tconv = lambda t, k: F.conv_transpose2d(t, w(k), b(k))
    up = lambda t, k: F.conv_transpose2d(t, w(k), b(k), stride=2)
    act = lambda t, k: _act(t, sd, k, activation)

    x = F.pad(x, (2, 2, 2, 2), mode="reflect")                     # UtNet.py:27,98
    skips = []
    for lvl in range(1, 5):                                         # UtNet.py:99-102
        if lvl > 1:
            x = F.max_pool2d(x, 2)
        x = act(conv(x, f"convs{lvl}.0"), f"convs{lvl}.1")
        x = act(conv(x, f"convs{lvl}.2"), f"convs{lvl}.3")
        skips.append(x)
    x = F.max_pool2d(x, 2)
    x = act(conv(x, "bottom.0"), "bottom.1")
    x = act(tconv(x, "bottom.2"), "bottom.3")                      # UtNet.py:51-57
    for lvl in range(1, 5):                                         # UtNet.py:103-107
        x = torch.cat([up(x, f"up{lvl}"), skips[4 - lvl]], dim=1)
        x = act(tconv(x, f"tconvs{lvl}.0"), f"tconvs{lvl}.1")
        x = act(tconv(x, f"tconvs{lvl}.2"), f"tconvs{lvl}.3")
    x = conv(x, "tconvs4.4")
    return x[:, :, 2:-2, 2:-2]                                      # ZeroPad2d(-2), UtNet.py:88,108


def unet_forward(sd, x: torch.Tensor, find_noise: bool = False) -> torch.Tensor:
    """UNet.forward (ThirdPartyNets.py:154-169), BatchNorm in eval mode."""

    def dconv(t, p):
        for cv, bn in ((".0", ".1"), (".3", ".4")):
            t = F.conv2d(t, sd[p + cv + ".weight"], sd[p + cv + ".bias"], padding=1)
            t = F.batch_norm(t, sd[p + bn + ".running_mean"], sd[p + bn + ".running_var"],
                             sd[p + bn + ".weight"], sd[p + bn + ".bias"], False, 0.0, 1e-5)
            t = F.relu(t)
        return t

    y = x
    xs = [dconv(x, "inc.conv.conv")]
    for i in range(1, 5):
        xs.append(dconv(F.max_pool2d(xs[-1], 2), f"down{i}.mpconv.1.conv"))
    t = xs[4]
    for i in range(1, 5):
        skip = xs[4 - i]
        t = F.conv_transpose2d(t, sd[f"up{i}.up.weight"], sd[f"up{i}.up.bias"], stride=2)
        dy, dx = skip.shape[2] - t.shape[2], skip.shape[3] - t.shape[3]
        t = F.pad(t, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))  # ThirdPartyNets.py:114-118
        t = dconv(torch.cat([skip, t], dim=1), f"up{i}.conv.conv")     # [skip, up] order, :124
    t = F.conv2d(t, sd["outc.conv.weight"], sd["outc.conv.bias"])
    return y - torch.sigmoid(t) if find_noise else torch.sigmoid(t)


def utnet_flops(cs: int, funit: int = 64) -> float:
    """Algorithmic 2*MAC per crop (SURVEY §8d closed form; Conv: out pixels, ConvT: in pixels)."""
    F_ = funit
    p1 = cs // 2
    p2 = (p1 - 4) // 2
    p3 = (p2 - 4) // 2
    p4 = (p3 - 4) // 2
    F2 = F_ * F_
    t = (27 * F_ * (cs + 2) ** 2 + 9 * F2 * cs ** 2 + 18 * F2 * (p1 - 2) ** 2 + 36 * F2 * (p1 - 4) ** 2 +
         72 * F2 * (p2 - 2) ** 2 + 144 * F2 * (p2 - 4) ** 2 + 288 * F2 * (p3 - 2) ** 2 +
         576 * F2 * (p3 - 4) ** 2 + (1152 + 2304) * F2 * (p4 - 2) ** 2 + 512 * F2 * p4 ** 2 +
         1152 * F2 * (p3 - 4) ** 2 + 576 * F2 * (p3 - 2) ** 2 + 128 * F2 * p3 ** 2 +
         288 * F2 * (p2 - 4) ** 2 + 144 * F2 * (p2 - 2) ** 2 + 32 * F2 * p2 ** 2 + 72 * F2 * (p1 - 4) ** 2 +
         36 * F2 * (p1 - 2) ** 2 + 8 * F2 * p1 ** 2 + 18 * F2 * cs ** 2 + 9 * F2 * (cs + 2) ** 2 +
         3 * F_ * (cs + 4) ** 2)
    return 2.0 * t


def unet_flops(cs: int) -> float:
    return 256.02e9 * (cs / 512.0) ** 2
