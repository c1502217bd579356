"""Drop-in ``nn.Module`` classes for the two NIND denoisers, backed by the sm_100a C-ABI library.

They mirror the reference classes' constructor signatures, parameter/buffer names, shapes and
``forward`` contract so that ``load_state_dict(strict=True)`` of a reference checkpoint works and
``nn_common.Model.instantiate_model(network='UtNet')`` can resolve to them (see ``register``):

  * ``UtNet(funit=64, activation='PReLU')``      — /root/reference/src/nind_denoise/networks/UtNet.py:13-109
  * ``UNet(n_channels=3, n_classes=3, funit=64, find_noise=False)``
                                                  — .../networks/ThirdPartyNets.py:138-169

The modules keep ordinary fp32 ``nn.Parameter``s (the master copy, what optimisers/state_dict see);
``forward`` hands them to ``nind_net_create`` which packs bf16 tap-major copies for the kernels, and
re-packs when a parameter's version counter changes.  Inference only (no autograd graph); CUDA
only — a CPU tensor raises, there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _capi

_ACTS = {"PReLU": nn.PReLU, "ELU": nn.ELU, "Hardswish": nn.Hardswish}


class _NativeNet(nn.Module):
    """Shared plumbing: lazily created native handle, re-pack on parameter change."""

    _arch = None

    def _init_native(self):
        self._handle = None
        self._packed_key = None
        self._options = {}

    # the handle is process-local state: never pickle/copy it (torch.save(model) must work, nn_common.py:73)
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_handle"] = None
        d["_packed_key"] = None
        return d

    def _drop_handle(self):
        h = getattr(self, "_handle", None)
        if h:
            try:
                _capi.lib().nind_net_destroy(h)
            except Exception:
                pass
        self._handle = None
        self._packed_key = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:  # interpreter shutdown: torch's module machinery may already be gone
            pass

    def _state_key(self):
        # storage identity + in-place version counter of every parameter and buffer: changes on
        # load_state_dict (copy_ bumps the version), optimiser steps, .to(device) (new storage)
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def set_option(self, key: str, value: int):
        """Tuning knob forwarded to nind_set_option (e.g. 'n_tile_deep', 'max_ctas')."""
        self._options[key] = int(value)
        if self._handle:
            _capi.check(_capi.lib().nind_set_option(self._handle, key.encode(), int(value)))

    def native_handle(self):
        """Create / refresh the packed native network for the current parameters and device."""
        key = self._state_key()
        if self._handle is not None and key == self._packed_key:
            return self._handle
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError(f"{type(self).__name__} (nind_denoise_b200) runs on CUDA sm_100 devices only; "
                               f"parameters are on {dev}. There is no CPU fallback.")
        if self._handle is not None and dev != getattr(self, "_device", dev):
            # .to(another GPU): the handle's streams, events, arenas and scratch live on the old device
            self._drop_handle()
        lib = _capi.lib()
        with torch.cuda.device(dev):
            arr, n, keep = _capi.make_tensor_array(self.state_dict())
            if self._handle is None:
                h = C.c_void_p()
                _capi.check(lib.nind_net_create(self._arch, int(self._funit), _capi.NIND_ACT[self._activation],
                                                arr, n, C.byref(h)))
                self._handle = h
                for k, v in self._options.items():
                    _capi.check(lib.nind_set_option(self._handle, k.encode(), v))
            else:
                _capi.check(lib.nind_net_load(self._handle, arr, n))
            del keep
        self._packed_key = key
        self._device = dev
        return self._handle

    def _forward_native(self, x: torch.Tensor, flags: int = 0) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__} (nind_denoise_b200): input must be a CUDA tensor, got "
                               f"{x.device}. There is no CPU fallback.")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected input [B,3,H,W], got {tuple(x.shape)}")
        h = self.native_handle()
        if x.device != self._device:
            raise RuntimeError(f"input on {x.device} but module on {self._device}")
        xin = x.detach().float().contiguous()
        out = torch.empty_like(xin)
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _capi.check(_capi.lib().nind_net_forward_ex(h, xin.data_ptr(), out.data_ptr(), xin.shape[0], xin.shape[2],
                                                        xin.shape[3], flags, C.c_void_p(stream)))
        return out

    def denoise_batch(self, noisy_batch: torch.Tensor) -> torch.Tensor:
        """``Generator.denoise_batch`` of the reference (nn_common.py:198-199: ``model(x).clip(0, 1)``), the call
        its validation / test passes make (nn_train.py:51-93); the clamp runs inside the output head's epilogue."""
        return self._forward_native(noisy_batch, _capi.NIND_FWD_CLAMP01)

    def layer_times(self, x: torch.Tensor):
        """Per-layer (name, ms, flops) of one forward, measured with CUDA events (debug/profiling aid)."""
        lib = _capi.lib()
        h = self.native_handle()
        _capi.check(lib.nind_set_timing(h, 1))
        try:
            self._forward_native(x)
            n = C.c_int()
            names = (C.c_char_p * 128)()
            ms = (C.c_float * 128)()
            fl = (C.c_double * 128)()
            _capi.check(lib.nind_get_layer_times(h, 128, names, ms, fl, C.byref(n)))
            return [(names[i].decode(), ms[i], fl[i]) for i in range(n.value)]
        finally:
            _capi.check(lib.nind_set_timing(h, 0))


class UtNet(_NativeNet):
    """U-Net with transposed convolutions (reference: networks/UtNet.py:13-109)."""

    _arch = _capi.NIND_ARCH_UTNET

    def __init__(self, funit=64, activation="PReLU"):
        super().__init__()
        funit = int(funit)  # nn_common.py:124 passes "k=v" strings
        if activation not in _ACTS:
            # the reference calls exit() here (UtNet.py:26); raising is the library-friendly equivalent
            raise ValueError(f"UtNet: unknown activation function: {activation}")
        self._funit, self._activation = funit, activation

        def act():
            return _ACTS[activation]() if activation == "PReLU" else _ACTS[activation](inplace=True)

        def block(kinds, chans):
            mods = []
            for kind, (ci, co) in zip(kinds, chans):
                mods.append(nn.Conv2d(ci, co, 3) if kind == "c" else nn.ConvTranspose2d(ci, co, 3))
                mods.append(act())
            return mods

        f = funit
        self.pad = nn.ReflectionPad2d(2)
        enc_in = 3
        for lvl in range(1, 5):
            co = f << (lvl - 1)
            setattr(self, f"convs{lvl}", nn.Sequential(*block("cc", [(enc_in, co), (co, co)])))
            if lvl == 1:
                self.maxpool = nn.MaxPool2d(2)
            enc_in = co
        self.bottom = nn.Sequential(*block("ct", [(8 * f, 16 * f), (16 * f, 16 * f)]))
        width = 16 * f
        for lvl in range(1, 5):
            half = width // 2
            setattr(self, f"up{lvl}", nn.ConvTranspose2d(width, half, 2, stride=2))
            mods = block("tt", [(width, half), (half, half)])
            if lvl == 4:
                mods.append(nn.Conv2d(f, 3, 1))
            setattr(self, f"tconvs{lvl}", nn.Sequential(*mods))
            width = half
        self.unpad = nn.ZeroPad2d(-2)
        self._init_native()

    def forward(self, l):
        return self._forward_native(l)


def _double_conv(ci, co):
    return nn.Sequential(nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True),
                         nn.Conv2d(co, co, 3, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))


class _Wrap(nn.Module):
    """Holds a child under a fixed attribute name to reproduce the reference's nested key paths."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            setattr(self, k, v)


class UNet(_NativeNet):
    """Classic padded U-Net (reference: networks/ThirdPartyNets.py:62-169); BatchNorm in eval mode is
    folded into the convolution weights, so ``forward`` always behaves like ``model.eval()``."""

    _arch = _capi.NIND_ARCH_UNET

    def __init__(self, n_channels=3, n_classes=3, funit=64, find_noise=False):
        super().__init__()
        if int(n_channels) != 3 or int(n_classes) != 3:
            raise ValueError("UNet (nind_denoise_b200): only n_channels=3, n_classes=3 is implemented")
        self._funit, self._activation = 64, "PReLU"  # funit is ignored by the reference too (:139-150)
        self.inc = _Wrap(conv=_Wrap(conv=_double_conv(3, 64)))
        for i, (ci, co) in enumerate([(64, 128), (128, 256), (256, 512), (512, 512)], start=1):
            setattr(self, f"down{i}", _Wrap(mpconv=nn.Sequential(nn.MaxPool2d(2), _Wrap(conv=_double_conv(ci, co)))))
        for i, (ci, co) in enumerate([(1024, 256), (512, 128), (256, 64), (128, 64)], start=1):
            setattr(self, f"up{i}", _Wrap(up=nn.ConvTranspose2d(ci // 2, ci // 2, 2, stride=2),
                                          conv=_Wrap(conv=_double_conv(ci, co))))
        self.outc = _Wrap(conv=nn.Conv2d(64, 3, 1))
        self.find_noise = bool(find_noise) if not isinstance(find_noise, str) else find_noise == "True"
        self.sigmoid = nn.Sigmoid()
        self._init_native()

    def forward(self, x):
        y = self._forward_native(x)
        if self.find_noise:  # ThirdPartyNets.py:167-168
            return x - y
        return y

    def denoise_batch(self, noisy_batch: torch.Tensor) -> torch.Tensor:
        if self.find_noise:  # the clamp applies to y - sigmoid(x), which is formed here
            return self.forward(noisy_batch).clip(0, 1)
        return self._forward_native(noisy_batch, _capi.NIND_FWD_CLAMP01)


def register(nn_common_module=None):
    """Make ``nn_common.Model.instantiate_model(network='UtNet'|'UNet')`` build these classes: the
    reference looks the class up with ``globals()[network]`` inside nn_common (nn_common.py:131,137)."""
    if nn_common_module is None:
        import nn_common as nn_common_module  # the reference module, must be importable
    nn_common_module.UtNet = UtNet
    nn_common_module.UNet = UNet
    return nn_common_module
