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


# ------------------------------------------------------------------ round 2
REF_SRC = "/root/reference/src/nind_denoise"


def _import_reference_nn_common():
    import sys
    import types
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    for missing in ("piqa",):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except Exception:
                sys.modules[missing] = types.ModuleType(missing)
    if not hasattr(sys.modules["piqa"], "MS_SSIM"):
        sys.modules["piqa"].MS_SSIM = type("MS_SSIM", (torch.nn.Module,), {})
        sys.modules["piqa"].SSIM = type("SSIM", (torch.nn.Module,), {})
    import nn_common
    return nn_common


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree is only present in the build container")
def test_register_into_the_real_reference_factory(tmp_path):
    """The drop-in claim, against the reference's own code: after register(), the unmodified
    nn_common.Model.instantiate_model (nn_common.py:116-138) builds THESE classes — also 'UNet', which the fork's
    factory cannot resolve on its own — and loads a state_dict checkpoint into them."""
    nn_common = _import_reference_nn_common()
    ref_utnet = nn_common.UtNet
    try:
        nb.register(nn_common)
        m = nn_common.Model.instantiate_model(models_dpath=None, network="UtNet", device=None)
        assert type(m) is nb.UtNet and len(m.state_dict()) == 64
        m = nn_common.Model.instantiate_model(models_dpath=None, network="UNet", device=None)
        assert type(m) is nb.UNet and len(m.state_dict()) == 136
        # strparameters arrive as strings (nn_common.py:124)
        m = nn_common.Model.instantiate_model(models_dpath=None, network="UtNet", device=None,
                                              strparameters="funit=64,activation=Hardswish")
        assert m._activation == "Hardswish" and len(m.state_dict()) == 46
        # a .pt checkpoint written by the reference's own class loads strictly
        torch.manual_seed(0)
        ckpt = tmp_path / "generator_3.pt"
        torch.save(ref_utnet().state_dict(), ckpt)
        m = nn_common.Model.instantiate_model(models_dpath=None, model_path=str(ckpt), network="UtNet", device=None,
                                              keyword="generator")
        ref = on.init_state_dict("UtNet", seed=0)
        assert type(m) is nb.UtNet and all(torch.equal(m.state_dict()[k], ref[k]) for k in ref)
        with pytest.raises(RuntimeError, match="no CPU fallback"):   # built on the CPU it refuses to run
            m(torch.rand(1, 3, 120, 120))
    finally:
        nn_common.UtNet = ref_utnet
        if hasattr(nn_common, "UNet"):
            del nn_common.UNet


def test_complete_path_semantics(tmp_path):
    """cli.complete_path == Model.complete_path (nn_common.py:75-114)."""
    import json
    from nind_denoise_b200 import cli
    d = tmp_path / "models" / "runA"
    d.mkdir(parents=True)
    for n in ("generator_5.pt", "generator_40.pt", "discriminator_77.pt"):
        (d / n).write_bytes(b"")
    f = str(d / "generator_5.pt")
    ref = _import_reference_nn_common().Model.complete_path if os.path.isdir(REF_SRC) else None
    assert cli.complete_path(f, None, "generator") == f                                   # a file: as is
    assert cli.complete_path(str(d), None, "generator") == str(d / "generator_40.pt")     # highest with the keyword
    assert cli.complete_path(str(d), None, "discriminator") == str(d / "discriminator_77.pt")
    # a name under models_dpath recurses WITHOUT the keyword (nn_common.py:111): highest file of any kind
    assert cli.complete_path("runA", str(tmp_path / "models"), "generator") == str(d / "discriminator_77.pt")
    if ref:   # same answers from the reference's own function
        assert ref(f, None, "generator") == f
        assert ref(str(d), None, "generator") == str(d / "generator_40.pt")
        assert ref(str(d), None, "discriminator") == str(d / "discriminator_77.pt")
        assert ref("runA", str(tmp_path / "models"), "generator") == str(d / "discriminator_77.pt")
    json.dump({"best_epoch": {"validation_loss": 5}}, open(d / "trainres.json", "w"))
    assert cli.complete_path(str(d), None, "generator") == str(d / "generator_5.pt")      # trainres.json wins
    if ref:
        assert ref(str(d), None, "generator") == str(d / "generator_5.pt")
    # (with trainres.json present the reference's highest-number search raises on that file name; here files
    # without a number are skipped)
    assert cli.complete_path(str(d), None, "discriminator") == str(d / "discriminator_77.pt")
    with pytest.raises(SystemExit):
        cli.complete_path("nope", str(tmp_path / "models"), "generator")


def test_scoring_matches_definitions():
    """scoring.ssim / ms_ssim (piqa 1.3 defaults) against a direct, un-separated restatement of the formulas, and
    the testres.json bookkeeping (json_saver.py)."""
    import json
    import tempfile
    from scipy.signal import convolve2d
    from nind_denoise_b200 import scoring
    rng = np.random.default_rng(5)
    a = rng.random((3, 40, 52)).astype(np.float32)
    b = np.clip(a + rng.normal(0, 0.1, a.shape), 0, 1).astype(np.float32)
    g = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
    g2 = np.outer(g / g.sum(), g / g.sum())
    vals = []
    for c in range(3):
        f = lambda z: convolve2d(z.astype(np.float64), g2, mode="valid")
        mx, my = f(a[c]), f(b[c])
        sxx, syy, sxy = f(a[c] * a[c]) - mx * mx, f(b[c] * b[c]) - my * my, f(a[c] * b[c]) - mx * my
        cs = (2 * sxy + 0.03 ** 2) / (sxx + syy + 0.03 ** 2)
        vals.append(((2 * mx * my + 0.01 ** 2) / (mx * mx + my * my + 0.01 ** 2) * cs).mean())
    got = float(scoring.ssim(torch.from_numpy(a)[None], torch.from_numpy(b)[None])[0])
    assert abs(got - float(np.mean(vals))) < 1e-5
    x = torch.rand(1, 3, 170, 180)
    assert abs(float(scoring.ssim(x, x)[0]) - 1) < 1e-6 and abs(float(scoring.ms_ssim(x, x)[0]) - 1) < 1e-5
    y = (x + 0.1 * torch.randn_like(x)).clip(0, 1)
    m1, m2 = float(scoring.ms_ssim(x, y)[0]), float(scoring.ms_ssim(x, (x + 0.3 * torch.randn_like(x)).clip(0, 1))[0])
    assert 0 < m2 < m1 < 1
    with pytest.raises(RuntimeError):
        scoring.ms_ssim(x[..., :100, :100], y[..., :100, :100])       # needs >= 162 px (pt_losses.py:20-29)
    l = scoring.get_losses(x[0], y[0])
    assert set(l) == {"mse", "ssim", "msssim"} and abs(l["msssim"] - (1 - m1)) < 1e-6
    assert scoring.avg_listofdicts([{"mse": 1.0, "ssim": 0.5}, {"mse": 3.0, "ssim": 0.25}]) == {"mse": 2.0, "ssim": 0.375}
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "testres.json")
        scoring.add_test_results(p, 340, {"mse": 5.6e-4, "ssim": 0.1168, "msssim": 0.0397})
        scoring.add_test_results(p, 141, {"mse": 5.5e-4, "ssim": 0.1146, "msssim": 0.0392})
        scoring.add_test_results(p, 200, {"mse": 9.9e-4, "ssim": 0.2, "msssim": 0.1})
        d = json.load(open(p))
        assert d["best_epoch"] == {"test_mse": 141, "test_ssim": 141, "test_msssim": 141}
        assert d["best_val"]["test_mse"] == 5.5e-4 and set(d) == {"best_val", "best_epoch", "340", "141", "200"}


def test_baseline_file_selection():
    """dir_cli.sort_isos / get_baseline_fpath == dataset_torch_3.sortISOs / get_baseline_fpath (:37-96)."""
    import tempfile
    from nind_denoise_b200 import dir_cli
    assert dir_cli.sort_isos(["ISO6400", "ISO200", "ISOH1", "ISO800"]) == (["ISO200"], ["ISO800", "ISO6400", "ISOH1"])
    assert dir_cli.sort_isos(["ISO200-2", "ISO200", "ISO3200"])[0] == ["ISO200", "ISO200-2"]
    assert dir_cli.sort_isos(["GT1", "a", "b"]) == (["GT1"], ["a", "b"])
    assert dir_cli.sort_isos(["b", "a", "c"]) == (["a"], ["b", "c"])
    with tempfile.TemporaryDirectory() as td:
        for n in ("NIND_x_ISO800.png", "NIND_x_ISO200.png", "NIND_x_ISOH2.png", "notes.txt"):
            open(os.path.join(td, n), "w").close()
        assert dir_cli.get_baseline_fpath(td) == os.path.join(td, "NIND_x_ISO200.png")
