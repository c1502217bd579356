"""CPU: the C-ABI library loads and exports what include/nind_b200.h declares; host-side geometry
(crop table, band rows, sharding) is bit-exact against the oracle; the nn.Module mirrors keep the
reference's state_dict layout and refuse to run without CUDA.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import nind_denoise_b200 as nb
from nind_denoise_b200 import _build, _capi
from oracle import geometry as og
from oracle import nets as on

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nind_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nind_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ctypes.CDLL(_build.build())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)


def test_crop_table_matches_oracle():
    sweep = [(W, H, cs, ucs, ol) for (W, H) in [(6000, 4000), (5999, 3999), (4000, 6000), (520, 504), (701, 333)]
             for (cs, ucs) in [(504, 480), (248, 224), (120, 96), (264, 224)] for ol in (0, 6, 16)
             if ucs <= min(W, H) or True]
    sweep += [(8256, 5504, 512, 384, 6), (8256, 5504, 440, 320, 6), (6000, 4000, 1016, 992, 32), (90, 70, 41, 28, 4)]
    for (W, H, cs, ucs, ol) in sweep:
        t = nb.crop_table(W, H, cs, ucs, ol)
        g = og.crop_grid(W, H, cs, ucs, ol)
        assert t.shape == (g.size, 8)
        assert np.array_equal(t, og.crop_table(g)), (W, H, cs, ucs, ol)
        assert nb.n_crops(W, H, cs, ucs, ol) == g.size


def test_golden_tables_through_c_abi(golden_geometry):
    G = golden_geometry
    for gi in range(7):
        W, H, cs, ucs, ol = (int(v) for v in G[f"g{gi}_params"])
        assert np.array_equal(nb.crop_table(W, H, cs, ucs, ol)[:, 2:], G[f"g{gi}_table"])


def test_illegal_geometry_is_an_error():
    with pytest.raises(_capi.NindError):
        nb.crop_table(100, 100, 40, 28, 28)  # stride 0
    with pytest.raises(_capi.NindError):
        nb.crop_table(100, 100, 28, 40, 4)  # ucs > cs


def test_band_rows_and_shards():
    W, H, cs, ucs, ol = 6000, 4000, 504, 480, 6
    n = nb.n_crops(W, H, cs, ucs, ol)
    for world in (1, 2, 4, 8):
        rs = nb.shard_ranges(n, world)
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(b - a for a, b in rs) == -(-n // world)
        covered = np.zeros(H, bool)
        for a, b in rs:
            if b > a:
                y0, y1 = _capi.band_rows(W, H, cs, ucs, ol, a, b)
                t = og.crop_table(og.crop_grid(W, H, cs, ucs, ol))
                assert y0 == t[a, 7] and y1 == min(H, t[b - 1, 7] + t[b - 1, 5] - t[b - 1, 3])
                covered[y0:y1] = True
        assert covered.all()


def test_module_state_dict_layout_matches_reference():
    for cls, net, kw in [(nb.UtNet, "UtNet", {}), (nb.UNet, "UNet", {})]:
        torch.manual_seed(0)
        m = cls(**kw)
        ref = on.init_state_dict(net, seed=0)
        sd = m.state_dict()
        assert list(sd.keys()) == list(ref.keys())
        for k in ref:
            assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype
            assert torch.equal(sd[k], ref[k]), k  # same init stream as the reference class
        m.load_state_dict(ref, strict=True)
    for act, n in (("ELU", 46), ("Hardswish", 46)):
        assert len(nb.UtNet(activation=act).state_dict()) == n
    assert nb.UtNet(funit="64")._funit == 64  # nn_common passes strings
    with pytest.raises(ValueError):
        nb.UtNet(activation="Swish")


def test_no_cpu_fallback():
    m = nb.UtNet()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 120, 120))
    with pytest.raises(RuntimeError):
        nb.denoise_tiled(torch.rand(3, 300, 300), m, 120, 96, 6)


def test_register_injects_into_factory_namespace():
    import types
    fake = types.ModuleType("nn_common")
    nb.register(fake)
    assert fake.UtNet is nb.UtNet and fake.UNet is nb.UNet


def test_module_pickles_without_native_handle(tmp_path):
    m = nb.UNet()
    torch.save(m, tmp_path / "m.pth")  # nn_common.py:73 saves whole modules as .pth
    m2 = torch.load(tmp_path / "m.pth", weights_only=False)
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())


def test_rows_needed_covers_mirror_padding():
    from nind_denoise_b200.tiler import rows_needed
    for (W, H, cs, ucs, ol) in [(1500, 1100, 248, 224, 6), (1500, 1100, 120, 96, 6), (6000, 4000, 504, 480, 6),
                                (700, 333, 248, 224, 6)]:
        g = og.crop_grid(W, H, cs, ucs, ol)
        t = og.crop_table(g)
        for world in (1, 3, 8):
            for a, b in nb.shard_ranges(g.size, world):
                if b <= a:
                    continue
                r0, r1 = rows_needed(W, H, cs, ucs, ol, a, b)
                used = np.concatenate([og._sym(np.arange(t[i, 1], t[i, 1] + cs), H) for i in range(a, b)])
                assert used.min() >= r0 and used.max() < r1, (W, H, cs, world, a, b)


def test_compute_entry_points_fail_loudly_without_a_gpu():
    """No GPU in this container: creating a network must return an error, never fall back to the CPU."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from oracle import nets as on_
    sd = on_.init_state_dict("UtNet", seed=0)
    arr, n, keep = _capi.make_tensor_array(sd)
    h = ctypes.c_void_p()
    rc = _capi.lib().nind_net_create(0, 64, 0, arr, n, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"cuda" in _capi.lib().nind_last_error().lower() or b"CUDA" in _capi.lib().nind_last_error()
    with pytest.raises(_capi.NindError):
        _capi.device_info()


def test_dir_cli_file_conventions(tmp_path):
    """dir_cli: listing, output naming (denoise_dir.py:84-85) and scoring helpers need no GPU."""
    from nind_denoise_b200 import dir_cli
    for n in ("b.png", "a.tif", "c.jpg", "notes.txt", "clean.tif"):
        (tmp_path / n).write_bytes(b"")
    ins = dir_cli.list_images(str(tmp_path), skip=[str(tmp_path / "clean.tif")])
    assert [os.path.basename(p) for p in ins] == ["a.tif", "b.png", "c.jpg"]
    assert dir_cli.out_path_for(ins[2], "/out") == "/out/c.jpg.tif"
    assert dir_cli.out_path_for(ins[0], "/out") == "/out/a.tif"
    a = torch.rand(3, 8, 9)
    assert dir_cli.losses(a, a)["mse"] == 0.0
    l = dir_cli.losses(a, (a + 0.1))
    assert abs(l["psnr"] - (-10 * np.log10(l["mse"]))) < 1e-4
