#!/usr/bin/env python
"""Drop-in for the reference's tiled-inference script:

    python -m nind_denoise_b200.cli --network UtNet --model_path generator_650.pt \\
           --input in_s1.tif --output out_s1_denoised.tiff

Same flags and file conventions as /root/reference/src/nind_denoise/denoise_image.py:181-200 (so
``src/denoise.py:430-436`` can spawn it unchanged); the crop loop runs through
``nind_tiled_denoise_host`` on a B200 instead of the per-crop Python loop.

File I/O restates, with cv2 only, what the reference helpers do:
  * read   common/libs/np_imgops.py:12-29  (BGR->RGB, HWC->CHW, u8/255, u16/65535, float passthrough)
  * write  common/libs/pt_helpers.py:22-40 (.jpg 8-bit clip; .png/.tif clip(0,1)*65535 round u16;
                                            .tiff unclamped float32)
EXIF copying (piexif / exiv2 in the reference, :272-279) is skipped unless those modules are present.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time

import numpy as np
import torch

CS_UNET, UCS_UNET = 440, 320      # denoise_image.py:40
CS_UTNET, UCS_UTNET = 504, 480    # denoise_image.py:41
CS_UNK, UCS_UNK = 512, 448        # denoise_image.py:42


def img_path_to_np_flt(fpath: str) -> np.ndarray:
    import cv2

    if not os.path.isfile(fpath):
        raise FileNotFoundError(fpath)
    img = cv2.imread(fpath, flags=cv2.IMREAD_COLOR + cv2.IMREAD_ANYDEPTH)
    if img is None:
        raise IOError(f"cannot decode {fpath}")
    rgb = cv2.cvtColor(img, cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    if rgb.dtype == np.float32:
        return np.ascontiguousarray(rgb)
    if rgb.dtype == np.uint8:
        return rgb.astype(np.float32) / 255
    if rgb.dtype == np.uint16:
        return rgb.astype(np.float32) / 65535
    raise TypeError(f"{fpath} has unknown pixel format {rgb.dtype}")


def tensor_to_imgfile(t: torch.Tensor, path: str) -> None:
    import cv2

    ext = path[-4:].lower()
    if ext in (".jpg", "jpeg"):
        arr = (t.clip(0, 1) * 255).add(0.5).clamp(0, 255).byte().cpu().numpy().transpose(1, 2, 0)
        cv2.imwrite(path, cv2.cvtColor(arr, cv2.COLOR_RGB2BGR))
    elif ext in (".png", ".tif"):
        arr = (t.clip(0, 1) * 65535).round().cpu().numpy().astype(np.uint16).transpose(1, 2, 0)
        cv2.imwrite(path, cv2.cvtColor(arr, cv2.COLOR_RGB2BGR))
    elif ext == "tiff":
        arr = t.cpu().numpy().astype(np.float32).transpose(1, 2, 0)
        cv2.imwrite(path, cv2.cvtColor(arr, cv2.COLOR_RGB2BGR))
    else:
        raise NotImplementedError(f"Extension in {path}")


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


def load_model(args, device):
    """nn_common.Model.instantiate_model (nn_common.py:116-138) restricted to the two supported classes."""
    import nind_denoise_b200 as nb

    params = {}
    if args.model_parameters:
        params.update(dict(p.split("=") for p in args.model_parameters.split(",")))
    classes = {"UtNet": nb.UtNet, "UNet": nb.UNet}
    if args.g_network not in classes:
        sys.exit(f"network {args.g_network} is not available in nind_denoise_b200 (UtNet, UNet)")
    path = args.model_path
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


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--cs", type=int)
    ap.add_argument("--ucs", type=int)
    ap.add_argument("-ol", "--overlap", default=6, type=int)
    ap.add_argument("-i", "--input", default="in.jpg", type=str)
    ap.add_argument("-o", "--output", type=str)
    ap.add_argument("-b", "--batch_size", type=int, default=0, help="crops per forward (0 = auto)")
    ap.add_argument("--debug", action="store_true")
    ap.add_argument("--exif_method", default="noexif", type=str)
    ap.add_argument("--g_network", "--network", "--arch", type=str)
    ap.add_argument("--model_path", required=True)
    ap.add_argument("--model_parameters", type=str)
    ap.add_argument("--max_subpixels", type=int)
    ap.add_argument("--whole_image", action="store_true")
    ap.add_argument("--pad", type=int)
    ap.add_argument("--models_dpath")
    args, _ = ap.parse_known_args(argv)
    autodetect_network_cs_ucs(args)
    if args.max_subpixels is not None and 3 * args.cs * args.cs > args.max_subpixels:
        sys.exit(f"denoise_image.py: crop of 3x{args.cs}x{args.cs} > {args.max_subpixels=} for {args.input=}; aborting")
    if not torch.cuda.is_available():
        sys.exit("nind_denoise_b200 needs a CUDA sm_100 device (no CPU fallback)")
    import nind_denoise_b200 as nb

    device = torch.device("cuda")
    if args.model_parameters is None and "activation" in args.model_path:  # denoise_image.py:222-225
        args.model_parameters = f"activation={args.model_path.split('activation')[-1].split('_')[1].split('_')[0]}"
    model = load_model(args, device)
    if args.output is None:
        root, leaf = os.path.split(args.model_path)
        os.makedirs(os.path.join(root, "test", "denoised_images"), exist_ok=True)
        args.output = os.path.join(root, "test", "denoised_images", f"{os.path.basename(args.input)}_{leaf}.tif")
    img = torch.from_numpy(img_path_to_np_flt(args.input))
    start = time.time()
    if args.whole_image:  # one forward over the mirror-padded image (denoise_image.py:91-97,110-128)
        if not args.pad:
            print("OneImageDS: Warning: you should really consider (pad>0)")
        out = nb.denoise_whole_image(img.to(device), model, args.pad or 0).cpu()
    else:
        out = nb.denoise_tiled_host(img.pin_memory(), model, args.cs, args.ucs, args.overlap,
                                    batch=args.batch_size or None)
    tensor_to_imgfile(out, args.output)
    print(f"Wrote denoised image to {args.output}")
    print("Elapsed time: " + str(time.time() - start) + " seconds")
    return 0


if __name__ == "__main__":
    sys.exit(main())
