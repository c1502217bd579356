#!/usr/bin/env python
"""Denoise every image of a directory in ONE process:

    python -m nind_denoise_b200.dir_cli --noisy_dir shots/ --result_dir out/ \\
           --network UtNet --model_path generator_650.pt [--baseline clean.tif]

The reference's /root/reference/src/nind_denoise/denoise_dir.py:76-103 spawns one `denoise_image.py`
process per image (model load + CUDA start-up each time).  Here the images stream through
``nind_tiled_denoise_host_async`` — image k+1's upload overlaps image k's compute and download — while
a small thread pool decodes the next files and encodes the finished ones.  File conventions are the
reference's (`.jpg` inputs are written as `<name>.jpg.tif`, denoise_dir.py:84-85; `--skip_existing`).

Scoring: with ``--baseline`` the MSE / PSNR of every output against that clean image is printed and the
averages returned (the reference also reports SSIM / MS-SSIM through `piqa`, pt_helpers.get_losses; that
package is not a dependency here, so those two are only added when it is importable).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence

import torch

from .cli import autodetect_network_cs_ucs, img_path_to_np_flt, load_model, tensor_to_imgfile

IMG_EXT = (".tif", ".tiff", ".png", ".jpg", ".jpeg")


def list_images(noisy_dir: str, skip: Sequence[str] = ()) -> List[str]:
    """Image files of ``noisy_dir`` in sorted order (the reference iterates os.listdir, denoise_dir.py:78)."""
    skip = {os.path.abspath(s) for s in skip if s}
    names = sorted(n for n in os.listdir(noisy_dir) if n.lower().endswith(IMG_EXT))
    return [os.path.join(noisy_dir, n) for n in names if os.path.abspath(os.path.join(noisy_dir, n)) not in skip]


def out_path_for(in_path: str, result_dir: str) -> str:
    out = os.path.join(result_dir, os.path.basename(in_path))
    return out + ".tif" if out.endswith("jpg") else out  # denoise_dir.py:84-85


def losses(clean: torch.Tensor, out: torch.Tensor) -> Dict[str, float]:
    """MSE / PSNR on 0..1 data (+ SSIM when piqa is available), cf. pt_helpers.get_losses."""
    a, b = clean.clamp(0, 1), out.clamp(0, 1)
    mse = float(((a - b) ** 2).mean())
    res = {"mse": mse, "psnr": float("inf") if mse == 0 else -10.0 * float(torch.log10(torch.tensor(mse)))}
    try:
        import piqa  # noqa: F401

        res["ssim"] = float(piqa.SSIM()(a[None], b[None]))
    except Exception:
        pass
    return res


def _load(path: str) -> torch.Tensor:
    t = torch.from_numpy(img_path_to_np_flt(path))
    return t.pin_memory() if torch.cuda.is_available() else t


def denoise_dir(in_paths: Sequence[str], out_paths: Sequence[str], model, cs: int, ucs: int, ol: int = 6,
                batch: Optional[int] = None, baseline: Optional[torch.Tensor] = None, group: int = 2,
                io_threads: int = 4, verbose: bool = True) -> List[Optional[Dict[str, float]]]:
    """Stream ``in_paths`` through the two-slot host pipeline, ``group`` images per synchronisation,
    decoding the next group and encoding the previous one on ``io_threads`` worker threads meanwhile."""
    import nind_denoise_b200 as nb

    assert len(in_paths) == len(out_paths)
    scores: List[Optional[Dict[str, float]]] = [None] * len(in_paths)
    groups = [list(range(i, min(len(in_paths), i + group))) for i in range(0, len(in_paths), group)]
    with ThreadPoolExecutor(max_workers=io_threads) as pool:
        pending = [pool.submit(_load, in_paths[i]) for i in groups[0]] if groups else []
        writes = []
        for gi, idxs in enumerate(groups):
            imgs = [f.result() for f in pending]
            pending = [pool.submit(_load, in_paths[i]) for i in groups[gi + 1]] if gi + 1 < len(groups) else []
            outs = nb.denoise_images_host(imgs, model, cs, ucs, ol, batch=batch)   # enqueue all, one sync
            for i, out in zip(idxs, outs):
                if baseline is not None and baseline.shape == out.shape:
                    scores[i] = losses(baseline, out)
                writes.append(pool.submit(tensor_to_imgfile, out, out_paths[i]))
                if verbose:
                    print(f"in: {in_paths[i]}, out: {out_paths[i]}" + (f", {scores[i]}" if scores[i] else ""))
        for w in writes:
            w.result()
    return scores


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--noisy_dir", required=True, type=str)
    ap.add_argument("--g_network", "--network", type=str)
    ap.add_argument("--model_path", "--model_fpath", required=True)
    ap.add_argument("--model_parameters", default="", type=str)
    ap.add_argument("--result_dir", default=None, type=str)
    ap.add_argument("--no_scoring", action="store_true")
    ap.add_argument("--baseline", type=str, help="clean image to score against (skipped as an input)")
    ap.add_argument("--cs", type=int)
    ap.add_argument("--ucs", type=int)
    ap.add_argument("-ol", "--overlap", default=6, type=int)
    ap.add_argument("-b", "--batch_size", type=int, default=0)
    ap.add_argument("--skip_existing", action="store_true")
    ap.add_argument("--whole_image", action="store_true")
    args, _ = ap.parse_known_args(argv)
    autodetect_network_cs_ucs(args)
    if args.whole_image:
        sys.exit("--whole_image is not part of the tiled hot path of nind_denoise_b200")
    if not torch.cuda.is_available():
        sys.exit("nind_denoise_b200 needs a CUDA sm_100 device (no CPU fallback)")
    result_dir = args.result_dir or os.path.join(args.noisy_dir, "..", "denoised",
                                                 os.path.basename(os.path.dirname(os.path.abspath(args.model_path))))
    os.makedirs(result_dir, exist_ok=True)
    ins = list_images(args.noisy_dir, skip=[args.baseline])
    outs = [out_path_for(p, result_dir) for p in ins]
    if args.skip_existing:
        keep = [k for k, o in enumerate(outs) if not os.path.isfile(o)]
        ins, outs = [ins[k] for k in keep], [outs[k] for k in keep]
    model = load_model(args, torch.device("cuda"))
    clean = None
    if args.baseline and not args.no_scoring:
        clean = torch.from_numpy(img_path_to_np_flt(args.baseline))
    start = time.time()
    scores = denoise_dir(ins, outs, model, args.cs, args.ucs, args.overlap, batch=args.batch_size or None,
                         baseline=clean)
    done = [s for s in scores if s]
    if done:
        print({k: sum(s[k] for s in done) / len(done) for k in done[0]})
    print(f"Denoised {len(ins)} images in {time.time() - start:.2f} seconds")
    return 0


if __name__ == "__main__":
    sys.exit(main())
