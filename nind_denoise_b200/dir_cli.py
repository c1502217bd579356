#!/usr/bin/env python
"""Denoise every image of a directory (or of every image-set sub-directory) and score the results:

    python -m nind_denoise_b200.dir_cli --noisy_dir shots/ --result_dir out/ \\
           --network UtNet --model_path generator_650.pt [--gpus 8]

The reference's /root/reference/src/nind_denoise/denoise_dir.py:76-103 spawns one ``denoise_image.py`` process per
image (model load + CUDA start-up each time) and then reads every output back from disk to score it.  Here the
images stream through ``nind_tiled_denoise_host_async`` — image k+1's upload overlaps image k's compute and
download — while a thread pool decodes the next files and encodes the finished ones; with ``--gpus N`` there is
one such replica per GPU, all fed from ONE queue of files (BASELINE configs[4], throughput mode).

File conventions are the reference's: per set the lowest-ISO file is the clean baseline and is not denoised
(``dataset_torch_3.get_baseline_fpath``, :89-96; ``--baseline`` overrides it), ``.jpg`` inputs are written as
``<name>.jpg.tif`` (denoise_dir.py:84-85), ``--skip_existing``, ``--result_dir make_subdirs``.  Scoring:
``pt_helpers.get_losses`` (mse, 1-SSIM, 1-MS-SSIM — nind_denoise_b200/scoring.py) of every output against its
set's baseline, averaged per set and over sets, printed and recorded under ``test_*`` keys in ``testres.json``
(and ``trainres.json`` when it exists) next to the model, as denoise_dir.py:99-124 intends.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import scoring
from .cli import autodetect_network_cs_ucs, complete_path, img_path_to_np_flt, load_model, tensor_to_imgfile

IMG_EXT = (".tif", ".tiff", ".png", ".jpg", ".jpeg")


# ------------------------------------------------------------------------------ file conventions
def sort_isos(raw: Sequence[str]) -> Tuple[List[str], List[str]]:
    """``dataset_torch_3.sortISOs`` (:37-86): (base ISOs, other ISOs).  Names that are not all ``ISO<n>`` sort
    alphabetically with any ``GT*`` name as the base; ``ISO<n>-<rep>`` repeats of the lowest ISO are bases too."""
    raw = list(raw)
    if any(iso[:3] != "ISO" for iso in raw):
        bisos = [i for i in raw if "GT" in i]
        isos = sorted(i for i in raw if "GT" not in i)
        if not bisos:
            bisos.append(isos.pop(0))
        return bisos, isos
    hisos, nums, dup = [], [], {}
    for iso in raw:
        if "H" in iso:
            hisos.append(iso)
        elif "-" in iso:
            val, _, rep = iso[3:].partition("-")
            nums.append(int(val))
            dup.setdefault(val, []).append(rep)
        else:
            nums.append(int(iso[3:]))
    base, *rest = sorted(nums)
    bases: List = [base]
    while rest and bases[0] == rest[0]:
        bases.append(str(rest.pop(0)) + "-" + dup[str(bases[0])].pop())
    for val, reps in dup.items():
        for rep in reps:
            rest[rest.index(int(val))] = val + "-" + rep
    return ["ISO" + str(b) for b in bases], ["ISO" + str(i) for i in rest] + sorted(hisos)


def get_baseline_fpath(dpath: str) -> Optional[str]:
    """``dataset_torch_3.get_baseline_fpath`` (:89-96): the file whose last ``_``-separated token (the ISO) sorts
    first.  None when the directory's names cannot be ordered that way."""
    names = [n for n in os.listdir(dpath) if n.lower().endswith(IMG_EXT)]
    if not names:
        return None
    by_iso = {n.split("_")[-1].split(".")[0]: n for n in names}
    try:
        bisos, _ = sort_isos(by_iso.keys())
        return os.path.join(dpath, by_iso[bisos[0]])
    except (ValueError, KeyError, IndexError):
        return None


def list_images(noisy_dir: str, skip: Sequence[str] = ()) -> List[str]:
    """Image files of ``noisy_dir`` in sorted order (the reference iterates os.listdir, denoise_dir.py:78)."""
    skip = {os.path.abspath(s) for s in skip if s}
    names = sorted(n for n in os.listdir(noisy_dir) if n.lower().endswith(IMG_EXT))
    return [os.path.join(noisy_dir, n) for n in names if os.path.abspath(os.path.join(noisy_dir, n)) not in skip]


def out_path_for(in_path: str, result_dir: str) -> str:
    out = os.path.join(result_dir, os.path.basename(in_path))
    return out + ".tif" if out.endswith("jpg") else out  # denoise_dir.py:84-85


def result_dir_for(args, model_path: str) -> str:
    """denoise_dir.py:55-59."""
    if args.result_dir == "make_subdirs":
        model_dname = os.path.basename(os.path.dirname(os.path.abspath(args.model_path)))
        return os.path.join(args.noisy_dir, "..", "denoised", model_dname, os.path.basename(os.path.normpath(args.noisy_dir)))
    return os.path.join(args.result_dir, os.path.abspath(model_path).split("/")[-2])


def losses(clean: torch.Tensor, out: torch.Tensor) -> Dict[str, float]:
    """``pt_helpers.get_losses`` (mse, ssim = 1 - SSIM, msssim = 1 - MS-SSIM) plus PSNR of the mse."""
    res = scoring.get_losses(clean, out)
    mse = res["mse"]
    res["psnr"] = float("inf") if mse == 0 else -10.0 * float(torch.log10(torch.tensor(mse)))
    return res


def _load(path: str) -> torch.Tensor:
    t = torch.from_numpy(img_path_to_np_flt(path))
    return t.pin_memory() if torch.cuda.is_available() else t


def _quantised(out: torch.Tensor, path: str) -> torch.Tensor:
    """What reading ``path`` back would give (the reference scores the FILE it wrote, pt_helpers.py:42-45)."""
    ext = path[-4:].lower()
    if ext in (".png", ".tif"):
        return (out.clip(0, 1) * 65535).round() / 65535
    if ext in (".jpg", "jpeg"):
        return (out.clip(0, 1) * 255).add(0.5).clamp(0, 255).floor() / 255
    return out


# ------------------------------------------------------------------------------ one replica
def denoise_dir(in_paths: Sequence[str], out_paths: Sequence[str], model, cs: int, ucs: int, ol: int = 6,
                batch: Optional[int] = None, baselines: Optional[Sequence[Optional[str]]] = None, group: int = 2,
                io_threads: int = 4, verbose: bool = True, baseline: Optional[torch.Tensor] = None,
                score_device=None) -> List[Optional[Dict[str, float]]]:
    """Stream ``in_paths`` through the two-slot host pipeline, ``group`` images per synchronisation,
    decoding the next group and encoding the previous one on ``io_threads`` worker threads meanwhile.
    ``baselines[i]`` (a file path) or ``baseline`` (one tensor for all) is the clean image output i is scored
    against."""
    import nind_denoise_b200 as nb

    assert len(in_paths) == len(out_paths)
    scores: List[Optional[Dict[str, float]]] = [None] * len(in_paths)
    groups = [list(range(i, min(len(in_paths), i + group))) for i in range(0, len(in_paths), group)]
    clean_cache: Dict[str, torch.Tensor] = {}

    def clean_for(i):
        if baseline is not None:
            return baseline
        p = baselines[i] if baselines else None
        if not p:
            return None
        if p not in clean_cache:
            clean_cache.clear()  # sets come one after another: keep one clean image at a time
            clean_cache[p] = torch.from_numpy(img_path_to_np_flt(p))
        return clean_cache[p]

    sdev = score_device if score_device is not None else getattr(model, "_device", None)
    with ThreadPoolExecutor(max_workers=io_threads) as pool:
        pending = [pool.submit(_load, in_paths[i]) for i in groups[0]] if groups else []
        writes = []
        for gi, idxs in enumerate(groups):
            imgs = [f.result() for f in pending]
            pending = [pool.submit(_load, in_paths[i]) for i in groups[gi + 1]] if gi + 1 < len(groups) else []
            outs = nb.denoise_images_host(imgs, model, cs, ucs, ol, batch=batch)   # enqueue all, one sync
            for i, out in zip(idxs, outs):
                clean = clean_for(i)
                if clean is not None and clean.shape == out.shape:
                    q = _quantised(out, out_paths[i])
                    if sdev is not None:
                        scores[i] = losses(clean.to(sdev), q.to(sdev))
                    else:
                        scores[i] = losses(clean, q)
                writes.append(pool.submit(tensor_to_imgfile, out, out_paths[i]))
                if verbose:
                    print(f"in: {in_paths[i]}, out: {out_paths[i]}" + (f", {scores[i]}" if scores[i] else ""))
        for w in writes:
            w.result()
    return scores


# ------------------------------------------------------------------------------ N replicas, one file queue
def _replica(rank: int, args, model_path: str, work, results):
    """One process per GPU: pulls (index, in, out, baseline) items from the shared queue until it is empty."""
    torch.cuda.set_device(rank)
    from .tiler import bind_host_to_gpu
    bind_host_to_gpu(rank)  # this replica's pinned buffers and decode threads next to its GPU's PCIe root port
    model = load_model(_Args(args, model_path), torch.device("cuda", rank))
    while True:
        items = []
        for _ in range(2):  # two images per synchronisation keep the two device slots busy
            try:
                items.append(work.get_nowait())
            except Exception:
                break
        if not items:
            break
        sc = denoise_dir([it[1] for it in items], [it[2] for it in items], model, args.cs, args.ucs, args.overlap,
                         batch=args.batch_size or None, baselines=[it[3] for it in items], verbose=False)
        for it, s in zip(items, sc):
            results.put((it[0], s))
    results.put((-1 - rank, None))  # this replica is done


class _Args:
    """argparse namespace with the resolved model path (picklable for the replica processes)."""

    def __init__(self, args, model_path):
        self.__dict__.update(vars(args))
        self.model_path = model_path


def denoise_dir_multi_gpu(items, args, model_path: str, gpus: int):
    """items: list of (index, in_path, out_path, baseline_path).  Returns {index: scores}."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    work, results = ctx.Queue(), ctx.Queue()
    for it in items:
        work.put(it)
    procs = [ctx.Process(target=_replica, args=(r, args, model_path, work, results)) for r in range(gpus)]
    for p in procs:
        p.start()
    import queue as _queue

    out, done = {}, 0
    while done < gpus:
        try:
            idx, sc = results.get(timeout=5)
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                for p in procs:
                    p.terminate()
                raise RuntimeError("a GPU replica died")
            continue
        if idx < 0:
            done += 1
        else:
            out[idx] = sc
    for p in procs:
        p.join()
        if p.exitcode != 0:
            raise RuntimeError(f"a GPU replica exited with code {p.exitcode}")
    return out


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--noisy_dir", required=True, type=str)
    ap.add_argument("--g_network", "--network", type=str)
    ap.add_argument("--model_path", "--model_fpath", required=True)
    ap.add_argument("--model_parameters", default="", type=str)
    ap.add_argument("--result_dir", default="../../results/NIND/test", type=str)
    ap.add_argument("--no_scoring", action="store_true")
    ap.add_argument("--baseline", type=str, help="clean image to score against (default: the lowest-ISO file of each set)")
    ap.add_argument("--cs", type=str)
    ap.add_argument("--ucs", type=str)
    ap.add_argument("-ol", "--overlap", default=6, type=int)
    ap.add_argument("-b", "--batch_size", type=int, default=0)
    ap.add_argument("--skip_existing", action="store_true")
    ap.add_argument("--whole_image", action="store_true")
    ap.add_argument("--models_dpath")
    ap.add_argument("--gpus", type=int, default=1, help="replicas, one per GPU, fed from one file queue (0 = all GPUs)")
    return ap


def main(argv=None) -> int:
    args, _ = build_parser().parse_known_args(argv)
    autodetect_network_cs_ucs(args)
    if args.whole_image:
        sys.exit("--whole_image is not part of the tiled hot path of nind_denoise_b200")
    if not torch.cuda.is_available():
        sys.exit("nind_denoise_b200 needs a CUDA sm_100 device (no CPU fallback)")
    model_path = complete_path(args.model_path, args.models_dpath, keyword="generator")
    result_dir = result_dir_for(args, model_path)
    os.makedirs(result_dir, exist_ok=True)
    # image sets: sub-directories, or the directory itself when it holds the images (denoise_dir.py:51-53)
    entries = sorted(os.listdir(args.noisy_dir))
    sets = ["."] if (not entries or os.path.isfile(os.path.join(args.noisy_dir, entries[0]))) else \
        [e for e in entries if os.path.isdir(os.path.join(args.noisy_dir, e))]
    items, set_of = [], []
    for si, aset in enumerate(sets):
        indir = os.path.join(args.noisy_dir, aset)
        base = args.baseline or (None if args.no_scoring else get_baseline_fpath(indir))
        for p in list_images(indir, skip=[base]):
            o = out_path_for(p, result_dir)
            if args.skip_existing and os.path.isfile(o):
                continue
            items.append((len(items), p, o, None if args.no_scoring else base))
            set_of.append(si)
    start = time.time()
    gpus = torch.cuda.device_count() if args.gpus == 0 else min(args.gpus, torch.cuda.device_count())
    if gpus > 1 and len(items) > 2:
        by_idx = denoise_dir_multi_gpu(items, args, model_path, gpus)
        scores = [by_idx.get(i) for i in range(len(items))]
    else:
        model = load_model(_Args(args, model_path), torch.device("cuda", torch.cuda.current_device()))
        scores = denoise_dir([it[1] for it in items], [it[2] for it in items], model, args.cs, args.ucs, args.overlap,
                             batch=args.batch_size or None, baselines=[it[3] for it in items])
    per_set = []
    for si in range(len(sets)):
        sc = [s for s, k in zip(scores, set_of) if k == si and s]
        if sc:
            per_set.append(scoring.avg_listofdicts(sc))
    total = scoring.avg_listofdicts(per_set)
    print(total)
    if total:
        leaf = os.path.basename(model_path)
        try:
            epoch = int(leaf.split("_")[1].split(".")[0])
        except (IndexError, ValueError) as e:
            print(f"Cannot determine epoch from model_path {model_path} ({e})")
            epoch = None
        res = {k: v for k, v in total.items() if k in ("mse", "ssim", "msssim")}
        mdir = os.path.dirname(os.path.abspath(model_path))
        if epoch is not None and os.path.isfile(os.path.join(mdir, "trainres.json")):
            scoring.add_test_results(os.path.join(mdir, "trainres.json"), epoch, res)
        scoring.add_test_results(os.path.join(mdir, "testres.json"), epoch, res)
    print(f"Denoised {len(items)} images in {time.time() - start:.2f} seconds")
    return 0


if __name__ == "__main__":
    sys.exit(main())
