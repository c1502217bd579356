#!/usr/bin/env python
"""Drop-in for the reference's tiled-inference script:

    python -m nind_denoise_b200.cli --network UtNet --model_path generator_650.pt \\
           --input in_s1.tif --output out_s1_denoised.tiff

Same flags and file conventions as /root/reference/src/nind_denoise/denoise_image.py:181-200 (so
``src/denoise.py:430-436`` can spawn it unchanged); the crop loop runs on a B200 instead of the per-crop Python
loop.  What happens to the pixels either side of the network also runs on the GPU:

  * read   common/libs/np_imgops.py:12-29: the file is decoded by cv2 as the reference does; the decoded
           interleaved u8 / u16 / f32 BGR array is uploaded as it is (half the PCIe bytes of the fp32 image for
           16-bit files) and ``nind_image_to_chw_f32`` does BGR->RGB, HWC->CHW and x/255 | x/65535 | passthrough;
  * write  common/libs/pt_helpers.py:22-40: ``nind_chw_f32_to_image`` does clip(0,1)*65535 round (.png/.tif),
           clip(0,1)*255+0.5 truncate (.jpg) or the unclamped float32 copy (.tiff), RGB->BGR and CHW->HWC; the
           interleaved result is downloaded and handed to cv2.imwrite.

``--model_path`` accepts what ``nn_common.Model.complete_path`` accepts (nn_common.py:75-114): a file, a directory
(best epoch from trainres.json, else the highest-numbered checkpoint) or a directory name under ``--models_dpath``.
EXIF (denoise_image.py:272-279): copied with exiv2 / piexif when those modules are importable, otherwise a warning
says that it was not.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

CS_UNET, UCS_UNET = 440, 320      # denoise_image.py:40
CS_UTNET, UCS_UTNET = 504, 480    # denoise_image.py:41
CS_UNK, UCS_UNK = 512, 448        # denoise_image.py:42

_PIX = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2}


# ------------------------------------------------------------------------------ file <-> tensor
def decode_image(fpath: str) -> np.ndarray:
    """The decode step of img_path_to_np_flt (np_imgops.py:16-19): cv2, colour, any depth -> [H,W,3] BGR."""
    import cv2

    if not os.path.isfile(fpath):
        raise FileNotFoundError(fpath)
    img = cv2.imread(fpath, flags=cv2.IMREAD_COLOR + cv2.IMREAD_ANYDEPTH)
    if img is None:
        raise IOError(f"cannot decode {fpath}")
    if img.dtype not in _PIX:
        raise TypeError(f"img_path_to_np_flt: Error: fpath={fpath} has unknown format ({img.dtype})")
    return np.ascontiguousarray(img)


def image_to_chw(hwc: torch.Tensor, bgr: bool = True) -> torch.Tensor:
    """[H,W,3] u8 / u16 / f32 CUDA tensor (cv2 channel order if ``bgr``) -> [3,H,W] fp32 RGB on the same device
    through ``nind_image_to_chw_f32`` (np_imgops.py:19-28)."""
    from . import _capi

    if not hwc.is_cuda or hwc.dim() != 3 or hwc.shape[2] != 3 or not hwc.is_contiguous():
        raise ValueError("image_to_chw expects a contiguous [H,W,3] CUDA tensor")
    dt = {torch.uint8: _capi.NIND_PIX_U8, torch.uint16: _capi.NIND_PIX_U16, torch.float32: _capi.NIND_PIX_F32}
    if hwc.dtype not in dt:
        raise TypeError(f"unsupported pixel type {hwc.dtype}")
    h, w = int(hwc.shape[0]), int(hwc.shape[1])
    out = torch.empty((3, h, w), dtype=torch.float32, device=hwc.device)
    with torch.cuda.device(hwc.device):
        _capi.check(_capi.lib().nind_image_to_chw_f32(hwc.data_ptr(), dt[hwc.dtype], h, w, int(bool(bgr)), out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def chw_to_image(chw: torch.Tensor, dtype: torch.dtype, bgr: bool = True) -> torch.Tensor:
    """[3,H,W] fp32 CUDA tensor -> [H,W,3] u8 / u16 / f32 (cv2 channel order if ``bgr``) through
    ``nind_chw_f32_to_image`` (pt_helpers.py:24-32: the clamp + quantisation of tensor_to_imgfile)."""
    from . import _capi

    if not chw.is_cuda or chw.dim() != 3 or chw.shape[0] != 3:
        raise ValueError("chw_to_image expects a [3,H,W] CUDA tensor")
    dt = {torch.uint8: _capi.NIND_PIX_U8, torch.uint16: _capi.NIND_PIX_U16, torch.float32: _capi.NIND_PIX_F32}
    chw = chw.detach().float().contiguous()
    h, w = int(chw.shape[1]), int(chw.shape[2])
    out = torch.empty((h, w, 3), dtype=dtype, device=chw.device)
    with torch.cuda.device(chw.device):
        _capi.check(_capi.lib().nind_chw_f32_to_image(chw.data_ptr(), h, w, dt[dtype], int(bool(bgr)), out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def img_path_to_np_flt(fpath: str) -> np.ndarray:
    """Host restatement of np_imgops.py:12-29 (CHW float32 RGB); the CLI itself converts on the GPU."""
    import cv2

    rgb = cv2.cvtColor(decode_image(fpath), cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    if rgb.dtype == np.float32:
        return np.ascontiguousarray(rgb)
    return rgb.astype(np.float32) / (255 if rgb.dtype == np.uint8 else 65535)


def load_image_cuda(fpath: str, device) -> torch.Tensor:
    """File -> [3,H,W] fp32 RGB on ``device``: decode on the host, convert on the GPU."""
    arr = decode_image(fpath)
    t = torch.from_numpy(arr.view(np.int16) if arr.dtype == np.uint16 else arr)  # torch.from_numpy has no uint16
    t = t.to(device, non_blocking=False)
    if arr.dtype == np.uint16:
        t = t.view(torch.uint16)
    return image_to_chw(t, bgr=True)


def output_pixel_type(path: str) -> torch.dtype:
    ext = path[-4:].lower()
    if ext in (".jpg", "jpeg"):
        return torch.uint8
    if ext in (".png", ".tif"):
        return torch.uint16
    if ext == "tiff":
        return torch.float32
    raise NotImplementedError(f"Extension in {path}")


def tensor_to_imgfile(t: torch.Tensor, path: str) -> None:
    """tensor_to_imgfile (pt_helpers.py:22-40) for a [3,H,W] fp32 tensor.  CUDA tensors are clamped / quantised /
    interleaved by ``nind_chw_f32_to_image`` and only the file-format bytes cross PCIe; CPU tensors take the
    equivalent torch expressions."""
    import cv2

    dtype = output_pixel_type(path)
    if t.is_cuda:
        q = chw_to_image(t, dtype, bgr=True)
        arr = (q.view(torch.int16) if dtype == torch.uint16 else q).cpu().numpy()
        arr = arr.view(np.uint16) if dtype == torch.uint16 else arr
    else:
        t = t.detach().float()
        if dtype == torch.uint8:
            rgb = (t.clip(0, 1) * 255).add(0.5).clamp(0, 255).byte().numpy()
        elif dtype == torch.uint16:
            rgb = (t.clip(0, 1) * 65535).round().numpy().astype(np.uint16)
        else:
            rgb = t.numpy().astype(np.float32)
        arr = cv2.cvtColor(np.ascontiguousarray(rgb.transpose(1, 2, 0)), cv2.COLOR_RGB2BGR)
    if not cv2.imwrite(path, arr):
        raise IOError(f"cannot write {path}")


# ------------------------------------------------------------------------------ model loading
def complete_path(path: str, models_dpath=None, keyword: str = "") -> str:
    """``Model.complete_path`` (nn_common.py:75-114): a file is returned as is; a directory resolves to the best
    epoch recorded in its trainres.json (generators only) or else its highest-numbered ``*_<n>.*`` file containing
    ``keyword``; a name that is a directory under ``models_dpath`` recurses — with the keyword DROPPED, as in the
    reference (its recursive call passes ``keyword`` in the ``models_dpath`` position, :111), so there the
    highest-numbered file of any kind wins.  An unknown path ends the process like the reference's ``exit``."""

    def find_highest(paths, model_t):
        best = [None, 0]
        for p in paths:
            try:
                curval = int(p.split("_")[-1].split(".")[0])
            except ValueError:
                continue  # the reference raises on e.g. 'trainres.json'; files without a number cannot win anyway
            if curval > best[1] and model_t in p:
                best = [p, curval]
        return best[0]

    def find_best(dpath, model_t):
        if model_t != "generator":
            return None
        res = os.path.join(dpath, "trainres.json")
        if not os.path.isfile(res):
            print(f"find_best did not find {res}")
            return None
        with open(res) as fp:
            best_epoch = json.load(fp)["best_epoch"]["validation_loss"]
        return os.path.join(dpath, f"generator_{best_epoch}.pt")

    if os.path.isfile(path):
        return path
    if os.path.isdir(path):
        best = find_best(path, keyword)
        if best is not None:
            return best
        highest = find_highest(os.listdir(path), keyword)
        if highest is None:
            sys.exit(f"Model path not found: no checkpoint in {path}")
        return os.path.join(path, highest)
    if models_dpath and os.path.isdir(os.path.join(models_dpath, path)):
        return complete_path(os.path.join(models_dpath, path), None, "")
    sys.exit("Model path not found: %s" % path)


def autodetect_network_cs_ucs(args) -> None:
    """denoise_image.py:59-79."""
    if args.g_network is None:
        low = args.model_path.lower()
        if "unet" in low:
            args.g_network = "UNet"
        elif "utnet" in low:
            args.g_network = "UtNet"
        else:
            sys.exit('Could not determine network architecture from path. Please specify a "--network" type '
                     "(typically UNet or UtNet)")
    if args.cs is None or args.ucs is None:
        if args.g_network == "UNet":
            args.cs, args.ucs = CS_UNET, UCS_UNET
        elif args.g_network == "UtNet":
            args.cs, args.ucs = CS_UTNET, UCS_UTNET
        else:
            args.cs, args.ucs = CS_UNK, UCS_UNK
        print(f"cs={args.cs}, ucs={args.ucs}")
    args.cs, args.ucs = int(args.cs), int(args.ucs)  # denoise_dir.py declares them as strings


def load_model(args, device):
    """nn_common.Model.instantiate_model (nn_common.py:116-138) restricted to the two supported classes."""
    import nind_denoise_b200 as nb

    params = {}
    if args.model_parameters:
        params.update(dict(p.split("=") for p in args.model_parameters.split(",")))
    classes = {"UtNet": nb.UtNet, "UNet": nb.UNet}
    if args.g_network not in classes:
        sys.exit(f"network {args.g_network} is not available in nind_denoise_b200 (UtNet, UNet)")
    path = complete_path(args.model_path, getattr(args, "models_dpath", None), keyword="generator")
    if path.endswith(".pth"):
        model = torch.load(path, map_location="cpu", weights_only=False)
        if type(model).__name__ in classes and not hasattr(model, "native_handle"):
            native = classes[type(model).__name__](**params)
            native.load_state_dict(model.state_dict())
            model = native
    elif path.endswith("pt"):
        model = classes[args.g_network](**params)
        model.load_state_dict(torch.load(path, map_location="cpu"))
    else:
        sys.exit(f"Error: unable to load invalid model path: {path}")
    return model.to(device).eval()


def copy_exif(src: str, dst: str, method: str) -> bool:
    """denoise_image.py:272-279.  (The reference's `.jpg` test reads ``args.output[:-4] == '.jpg'``, which is never
    true, so every method but 'noexif' goes through exiv2; the same order is kept here, with piexif as the
    fall-back for its intended case.)  Returns whether metadata was copied; warns when it could not be."""
    if method == "noexif":
        return False
    try:
        import exiv2

        s = exiv2.ImageFactory.open(src)
        s.readMetadata()
        d = exiv2.ImageFactory.open(dst)
        d.setExifData(s.exifData())
        d.writeMetadata()
        return True
    except ImportError:
        pass
    except Exception as e:  # unreadable metadata must not lose the denoised image
        print(f"warning: exiv2 could not copy EXIF from {src} to {dst}: {e}", file=sys.stderr)
        return False
    if dst.lower().endswith((".jpg", ".jpeg")) and method == "piexif":
        try:
            import piexif

            piexif.transplant(src, dst)
            return True
        except ImportError:
            pass
    print(f"warning: --exif_method {method}: neither the exiv2 nor the piexif module is importable; "
          f"EXIF was NOT copied to {dst} (pass --exif_method noexif to silence this)", file=sys.stderr)
    return False


def debug_dump(img: torch.Tensor, model, args) -> None:
    """--debug (denoise_image.py:260-266): per crop, the noisy crop, the network output and the trimmed, seam-halved
    tile as JPEGs under ./dbg, plus the geometry line.  Uses the stand-alone gather op and plain forwards."""
    import nind_denoise_b200 as nb

    os.makedirs("dbg", exist_ok=True)
    table = nb.crop_table(img.shape[2], img.shape[1], args.cs, args.ucs, args.overlap)
    for i in range(table.shape[0]):
        crop = nb.gather_crops(model, img, args.cs, args.ucs, args.overlap, i, i + 1)
        den = model(crop)[0]
        x0, y0, ux0, uy0, ux1, uy1, ax, ay = (int(v) for v in table[i])
        t = den[:, uy0:uy1, ux0:ux1].clone()
        ol = args.overlap
        if ax != 0:
            t[:, :, 0:ol] /= 2
        if ay != 0:
            t[:, 0:ol, :] /= 2
        if ax + args.ucs < img.shape[2] and ol:
            t[:, :, -ol:] /= 2
        if ay + args.ucs < img.shape[1] and ol:
            t[:, -ol:, :] /= 2
        tensor_to_imgfile(den, f"dbg/crop{i}_0_denoised.jpg")
        tensor_to_imgfile(t, f"dbg/crop{i}_0_tensimg.jpg")
        tensor_to_imgfile(crop[0], f"dbg/crop{i}_0_noisy.jpg")
        print(tuple(t.shape))
        print((ax, ay, (ux0, uy0, ux1, uy1)))


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--cs", type=int)
    ap.add_argument("--ucs", type=int)
    ap.add_argument("-ol", "--overlap", default=6, type=int)
    ap.add_argument("-i", "--input", default="in.jpg", type=str)
    ap.add_argument("-o", "--output", type=str)
    ap.add_argument("-b", "--batch_size", type=int, default=0, help="crops per forward (0 = auto)")
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--exif_method", default="piexif", type=str)
    ap.add_argument("--g_network", "--network", "--arch", type=str)
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--model_parameters", type=str)
    ap.add_argument("--max_subpixels", type=int)
    ap.add_argument("--whole_image", action="store_true")
    ap.add_argument("--pad", type=int)
    ap.add_argument("--models_dpath")
    return ap


def main(argv=None) -> int:
    args, _ = build_parser().parse_known_args(argv)
    autodetect_network_cs_ucs(args)
    if args.max_subpixels is not None and 3 * args.cs * args.cs > args.max_subpixels:
        sys.exit(f"denoise_image.py: crop of 3x{args.cs}x{args.cs} > {args.max_subpixels=} for {args.input=}; aborting")
    if not torch.cuda.is_available():
        sys.exit("nind_denoise_b200 needs a CUDA sm_100 device (no CPU fallback)")
    import nind_denoise_b200 as nb

    device = torch.device("cuda", torch.cuda.current_device())
    if args.model_parameters is None and "activation" in args.model_path:  # denoise_image.py:222-225
        args.model_parameters = f"activation={args.model_path.split('activation')[-1].split('_')[1].split('_')[0]}"
    model = load_model(args, device)
    if args.output is None:
        root, leaf = os.path.split(args.model_path)
        os.makedirs(os.path.join(root, "test", "denoised_images"), exist_ok=True)
        args.output = os.path.join(root, "test", "denoised_images", f"{os.path.basename(args.input)}_{leaf}.tif")
    img = load_image_cuda(args.input, device)
    start = time.time()
    if args.whole_image:  # one forward over the mirror-padded image (denoise_image.py:91-97,110-128)
        if not args.pad:
            print("OneImageDS: Warning: you should really consider (pad>0)")
        out = nb.denoise_whole_image(img, model, args.pad or 0)
    else:
        if args.debug:
            debug_dump(img, model, args)
        out = nb.denoise_tiled(img, model, args.cs, args.ucs, args.overlap, batch=args.batch_size or None)
    tensor_to_imgfile(out, args.output)
    print(f"Denoised image written to {args.output}")
    copy_exif(args.input, args.output, args.exif_method)
    print(f"Wrote denoised image to {args.output}")
    print("Elapsed time: " + str(time.time() - start) + " seconds")
    return 0


if __name__ == "__main__":
    sys.exit(main())
