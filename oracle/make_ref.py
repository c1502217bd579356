"""ORACLE (test infrastructure only).  Recipe for ``oracle/_ref/``: the UNMODIFIED reference implementation of
the hot path, taken from where it lies under /root/reference, so that ``bench.py --impl reference`` (and the
``cpu_baseline`` leg) can time the reference's own classes on the GPU box's host cores — that box has no
/root/reference, but ``oracle/_ref/`` travels with the snapshot (it is git-ignored, never committed: no reference
source enters the repository's history).

    python oracle/make_ref.py        # run in the build container; __graft_entry__.build() does it too

What is taken (paths relative to /root/reference/src/nind_denoise): ``networks/UtNet.py``,
``networks/ThirdPartyNets.py`` (the two network classes) and ``denoise_image.py`` (``OneImageDS``,
G1/G2) plus the helper modules those import (``nn_common.py``, ``common/libs/*.py``).  ``load()`` below imports them
with the same stub modules for the absent third-party packages that ``oracle/make_golden.py`` uses.
"""
from __future__ import annotations

import math
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/nind_denoise"
DST = os.path.join(ROOT, "oracle", "_ref", "nind_denoise")
FILES = ["networks/UtNet.py", "networks/ThirdPartyNets.py", "denoise_image.py", "nn_common.py",
         "common/libs/np_imgops.py", "common/libs/pt_helpers.py", "common/libs/pt_losses.py",
         "common/libs/utilities.py", "common/libs/json_saver.py", "common/libs/pt_ops.py",
         "configs/common_conf_default.yaml"]


def make() -> bool:
    """Copy the files; returns False (and leaves an existing copy alone) when /root/reference is absent."""
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "networks", "UtNet.py"))


def load():
    """Import the vendored reference: returns (UtNet class, UNet class, OneImageDS factory over an in-memory
    image).  Missing third-party modules the reference imports at top level (configargparse, exiv2, imageio, piqa,
    piexif) are stubbed — none of them is on the hot path."""
    import torch

    if DST not in sys.path:
        sys.path.insert(0, DST)
    for missing in ("configargparse", "exiv2", "imageio", "piqa", "piexif"):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                sys.modules[missing] = types.ModuleType(missing)
    if not hasattr(sys.modules["piqa"], "MS_SSIM"):  # pt_losses.py subclasses these at import time
        sys.modules["piqa"].MS_SSIM = type("MS_SSIM", (torch.nn.Module,), {})
        sys.modules["piqa"].SSIM = type("SSIM", (torch.nn.Module,), {})
    from networks.ThirdPartyNets import UNet
    from networks.UtNet import UtNet

    cwd = os.getcwd()
    os.chdir(DST)  # denoise_image reads configs/common_conf_default.yaml relative to the cwd at import
    try:
        import denoise_image as di
    finally:
        os.chdir(cwd)

    def dataset(img, cs, ucs, ol):
        """The reference's OneImageDS over an in-memory CHW float32 array (its constructor only reads files)."""
        ds = di.OneImageDS.__new__(di.OneImageDS)
        ds.inimg = img
        ds.width, ds.height = img.shape[2], img.shape[1]
        ds.whole_image = False
        ds.cs, ds.ucs, ds.ol = cs, ucs, ol
        ds.iperhl = math.ceil((ds.width - ucs) / (ucs - ol))
        ds.pad = int((cs - ucs) / 2)
        ipervl = math.ceil((ds.height - ucs) / (ucs - ol))
        ds.size = (ds.iperhl + 1) * (ipervl + 1)
        return ds

    return UtNet, UNet, dataset


if __name__ == "__main__":
    ok = make()
    print("oracle/_ref:", "ready" if ok and available() else "unavailable (no /root/reference and no previous copy)")
