"""ctypes binding of include/nind_b200.h.  There is no fallback: if the library is missing, import of
the compute entry points raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

NIND_ARCH_UTNET, NIND_ARCH_UNET = 0, 1
NIND_ACT = {"PReLU": 0, "ELU": 1, "Hardswish": 2}
NIND_FWD_CLAMP01 = 1
NIND_PIX_U8, NIND_PIX_U16, NIND_PIX_F32 = 0, 1, 2


class NindTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class NindCrop(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("ud_x0", C.c_int32), ("ud_y0", C.c_int32),
                ("ud_x1", C.c_int32), ("ud_y1", C.c_int32), ("start_x", C.c_int32), ("start_y", C.c_int32)]


class NindError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nind_b200 error {code}: {msg}")
        self.code = code


_lib = None

# symbol -> (restype, argtypes); every symbol include/nind_b200.h declares
SIGNATURES = {
    "nind_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "nind_net_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(NindTensor), C.c_int, C.POINTER(C.c_void_p)]),
    "nind_net_load": (C.c_int, [C.c_void_p, C.POINTER(NindTensor), C.c_int]),
    "nind_net_destroy": (None, [C.c_void_p]),
    "nind_net_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nind_net_forward_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "nind_net_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "nind_image_to_chw_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nind_chw_f32_to_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nind_crop_table": (C.c_int, [C.c_int] * 5 + [C.POINTER(NindCrop), C.POINTER(C.c_int)]),
    "nind_tiled_denoise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 +
                           [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "nind_tiled_denoise_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 9 +
                                [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
    "nind_plan_steps": (C.c_int, [C.c_void_p] + [C.c_int] * 8 + [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "nind_copy_planes": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong,
                                   C.c_void_p]),
    "nind_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]),
    "nind_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "nind_peer_close": (C.c_int, [C.c_void_p]),
    "nind_peer_free": (C.c_int, [C.c_void_p]),
    "nind_add_rows": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_longlong, C.c_void_p]),
    "nind_gather_crops": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_void_p]),
    "nind_stitch_crops": (C.c_int, [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.POINTER(C.c_int),
                                                                    C.POINTER(C.c_int), C.c_void_p]),
    "nind_band_rows": (C.c_int, [C.c_int] * 7 + [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nind_tiled_denoise_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "nind_tiled_denoise_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6),
    "nind_host_sync": (C.c_int, [C.c_void_p]),
    "nind_tiled_denoise_host_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 10
                                      + [C.POINTER(C.c_void_p)]),
    "nind_host_join": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nind_host_join_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "nind_host_rows_done": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]),
    "nind_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "nind_host_unregister": (C.c_int, [C.c_void_p]),
    "nind_kernel_launches": (C.c_int64, []),
    "nind_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "nind_get_layer_times": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float),
                                       C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "nind_get_layer_bytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "nind_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "nind_last_error": (C.c_char_p, []),
}


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) libnind_b200.so.  Raises if unavailable."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        alt = os.environ.get("NIND_LIB")  # development A/B runs: another build of the same ABI (tools/build_base.sh)
        if alt:
            path = alt
        elif _build.is_stale():
            try:
                _build.build()
            except Exception as e:  # no nvcc on this machine: use the prebuilt library if there is one
                if not os.path.exists(path):
                    raise RuntimeError(f"libnind_b200.so is missing and cannot be built: {e}") from e
        l = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise NindError(rc, lib().nind_last_error().decode("utf-8", "replace"))


def device_info():
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    check(lib().nind_device_info(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def crop_table(width: int, height: int, cs: int, ucs: int, ol: int):
    """int32 array [n, 8]: x0, y0, ud_x0, ud_y0, ud_x1, ud_y1, start_x, start_y (host-only, no GPU)."""
    import numpy as np

    n = C.c_int()
    check(lib().nind_crop_table(width, height, cs, ucs, ol, None, C.byref(n)))
    arr = (NindCrop * n.value)()
    check(lib().nind_crop_table(width, height, cs, ucs, ol, arr, C.byref(n)))
    return np.frombuffer(arr, dtype=np.int32).reshape(n.value, 8).copy()


def band_rows(width, height, cs, ucs, ol, crop_begin, crop_end):
    y0, y1 = C.c_int(), C.c_int()
    check(lib().nind_band_rows(width, height, cs, ucs, ol, crop_begin, crop_end, C.byref(y0), C.byref(y1)))
    return y0.value, y1.value


def make_tensor_array(state_dict):
    """state_dict (name -> contiguous fp32 torch tensor, any device) -> (NindTensor array, keepalive)."""
    items = [(k, v) for k, v in state_dict.items() if v.dtype.is_floating_point]
    arr = (NindTensor * len(items))()
    keep = []
    for i, (k, v) in enumerate(items):
        t = v.detach().float().contiguous()
        keep.append(t)
        nb = k.encode()
        keep.append(nb)
        arr[i].name = nb
        arr[i].data = t.data_ptr()
        arr[i].ndim = t.dim()
        for d in range(t.dim()):
            arr[i].shape[d] = t.shape[d]
    return arr, len(items), keep
