"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star): crop indexing and stitch geometry bit-exact; pixels
max abs error <= 2e-2 and PSNR difference <= 0.05 dB versus the fp32 reference on 0..1 data.  Because
default-initialised weights give a tiny output range (sigma_out ~ 0.007) every pixel test
also bounds the error RELATIVE to the output's standard deviation."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

import nind_denoise_b200 as nb
from nind_denoise_b200 import _capi
from oracle import geometry as og
from oracle import nets as on

pytestmark = pytest.mark.gpu

MAX_ABS = 2e-2          # north_star tolerance
MAX_REL_SIGMA = 0.25    # max abs error / sigma_out (bf16 activations through 23 layers)
PSNR_DIFF = 0.05


def dev():
    assert torch.cuda.is_available(), "GPU tests need CUDA"
    return torch.device("cuda:0")


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)


def check_pixels(got, ref, what):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref).max()
    sigma = ref.std()
    print(f"{what}: max abs {err:.3e}, sigma_out {sigma:.4f}, err/sigma {err / sigma:.3f}, "
          f"PSNR(got, ref) {psnr(got, ref):.1f} dB")
    assert np.isfinite(got).all()
    assert err <= MAX_ABS, what
    assert err <= MAX_REL_SIGMA * sigma, what


@pytest.fixture(scope="module")
def utnet():
    assert _capi.device_info()[0] == 10, "needs an sm_100 device"
    m = nb.UtNet().to(dev()).eval()
    m.load_state_dict(on.init_state_dict("UtNet", seed=0))
    return m


@pytest.fixture(scope="module")
def unet():
    m = nb.UNet().to(dev()).eval()
    m.load_state_dict(on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7))
    return m


# ------------------------------------------------------------------ geometry: bit-exact
def test_gather_crops_bit_exact(utnet, golden_geometry):
    G = golden_geometry
    for gi in range(7):
        W, H, cs, ucs, ol = (int(v) for v in G[f"g{gi}_params"])
        img = torch.from_numpy(G[f"g{gi}_img"]).to(dev())
        crops = nb.gather_crops(utnet, img, cs, ucs, ol).cpu().numpy()
        g = og.crop_grid(W, H, cs, ucs, ol)
        for k, i in enumerate(G[f"g{gi}_crop_idx"]):
            assert np.array_equal(crops[int(i)], G[f"g{gi}_crops"][k]), (gi, i)
        for i in range(g.size):
            assert np.array_equal(crops[i], og.gather_crop(G[f"g{gi}_img"], g, i)), (gi, i)


def test_stitch_bit_exact(utnet, golden_geometry):
    G = golden_geometry
    for gi in range(7):
        W, H, cs, ucs, ol = (int(v) for v in G[f"g{gi}_params"])
        img = G[f"g{gi}_img"]
        g = og.crop_grid(W, H, cs, ucs, ol)
        x = torch.from_numpy(np.stack([og.gather_crop(img, g, i) for i in range(g.size)]))
        ramp = torch.linspace(0.5, 1.5, cs).view(1, 1, 1, -1) * torch.linspace(1.25, 0.75, cs).view(1, 1, -1, 1)
        outs = (x * ramp + 0.125).to(dev())
        band, y0, y1 = nb.stitch_crops(outs, H, W, cs, ucs, ol)
        assert (y0, y1) == (0, H)
        assert np.array_equal(band.cpu().numpy(), G[f"g{gi}_stitched"]), gi
        # partial ranges add up to the whole (exactly here: the fake model makes <=2-term sums order-free
        # only away from 4-way corners, so compare with tolerance of one ulp)
        parts = np.zeros((3, H, W), np.float32)
        for a, b in nb.shard_ranges(g.size, 3):
            if b > a:
                bd, p0, p1 = nb.stitch_crops(outs[a:b], H, W, cs, ucs, ol, a, b)
                parts[:, p0:p1] += bd.cpu().numpy()
        assert np.abs(parts - G[f"g{gi}_stitched"]).max() <= 1e-6


def test_stitch_full_size_property(utnet):
    # BASELINE config 2 geometry at full size: constant-one crops must stitch to exactly 1 everywhere,
    # through the real gather (mirror pad) of a constant image.
    W, H, cs, ucs, ol = 6000, 4000, 504, 480, 6
    n = nb.n_crops(W, H, cs, ucs, ol)
    assert n == 117
    ones = torch.ones((n, 3, cs, cs), device=dev())
    band, y0, y1 = nb.stitch_crops(ones, H, W, cs, ucs, ol)
    assert (y0, y1) == (0, H) and bool((band == 1.0).all())
    # gather of a coordinate image: every crop pixel equals img[sym(y), sym(x)]
    yy = torch.arange(H, device=dev(), dtype=torch.float32).view(1, H, 1).expand(1, H, W)
    xx = torch.arange(W, device=dev(), dtype=torch.float32).view(1, 1, W).expand(1, H, W)
    img = torch.cat([yy, xx, yy * 0 + 7], 0).contiguous()
    t = nb.crop_table(W, H, cs, ucs, ol)
    for i in (0, 12, 58, 104, 116):
        c = nb.gather_crops(utnet, img, cs, ucs, ol, i, i + 1)[0].cpu().numpy()
        ys = og._sym(np.arange(t[i, 1], t[i, 1] + cs), H)
        xs = og._sym(np.arange(t[i, 0], t[i, 0] + cs), W)
        assert np.array_equal(c[0], np.broadcast_to(ys[:, None], (cs, cs)).astype(np.float32))
        assert np.array_equal(c[1], np.broadcast_to(xs[None, :], (cs, cs)).astype(np.float32))


# ------------------------------------------------------------------ networks: tolerance
def test_utnet_forward_config1(utnet, golden_networks):
    """BASELINE config 1 ('one 3x256x256 crop' -> nearest legal size 248) + cs 120."""
    for cs in (120, 248):
        torch.manual_seed(1)
        x = torch.rand(1, 3, cs, cs)
        y = utnet(x.to(dev())).cpu().numpy()[0]
        assert y.shape == (3, cs, cs)
        check_pixels(y, golden_networks[f"utnet_out_{cs}"], f"UtNet cs={cs} vs reference golden")


def test_utnet_forward_cs264_and_104(utnet):
    """SURVEY 8d config 1 'also run 264' (the next legal size above 256) and the smallest legal crop, against the
    oracle (the reference classes equal the oracle exactly, tests/test_oracle_golden.py)."""
    sd = on.init_state_dict("UtNet", seed=0)
    for cs in (264, 104):
        torch.manual_seed(cs)
        x = torch.rand(1, 3, cs, cs)
        with torch.no_grad():
            ref = on.utnet_forward(sd, x).numpy()[0]
        y = utnet(x.to(dev())).cpu().numpy()[0]
        check_pixels(y, ref, f"UtNet cs={cs} vs oracle")


LAYER_SHAPES = [
    # taps cin n_total B Hs Ws a bo epi [n_tile ws ctas cg flat pair pool]: the layer shapes of UtNet at small sizes
    "9 8 64 2 44 52 0 0 0",                      # first layer (8-channel hi/lo input)
    "9 64 64 2 42 50 0 0 0 0 -1 0 0 -1 1 1",     # convs1.2: pixel-pair mode + fused pool
    "9 64 64 2 44 52 0 0 2 0 -1 0 0 -1 1 0",     # tconvs4.2 + fused 1x1 head (pixel-pair mode)
    "9 64 64 2 44 52 0 0 0 0 -1 0 0 -1 0 0",     # the same layer in the plain 3x3 form
    "9 128 64 2 44 52 0 0 0",                    # tconvs4.0
    "9 64 128 2 40 40 0 0 0 0 -1 0 0 -1 0 1",    # convs2.0 .. with fused pool
    "9 128 128 2 40 40 0 0 0",
    "9 256 128 2 36 36 0 0 0",
    "9 128 256 2 30 30 0 0 0",
    "9 256 256 3 28 28 0 0 0",
    "9 512 256 2 28 28 0 0 0",
    "9 256 512 3 14 14 0 0 0",
    "9 512 512 4 14 14 0 0 0",                   # flat tiles
    "1 128 256 2 41 37 0 0 1",                   # 2x2/s2 up-convs: depth-to-space epilogue, odd sizes (flat tiles)
    "1 128 256 2 40 48 0 0 1",                   # the same layer on 16x8 tiles: TMA stores, one box per 4 rows
    "1 256 512 2 30 30 0 0 1",
    "1 512 1024 2 14 14 0 0 1",
    "1 1024 2048 2 6 6 0 0 1",
]


@pytest.mark.parametrize("shape", LAYER_SHAPES)
def test_each_layer_shape_against_a_naive_convolution(shape):
    """Per-layer parity (SURVEY 8c protocol step 2): every layer shape of the network through `igemm_kernel`, alone,
    against a plain CUDA-core fp32 convolution over the same bf16 inputs (tools/probe, built by
    __graft_entry__.build() from the library's own kernel objects): isolates a kernel bug from bf16 drift through
    23 layers.  The probe also checks that nothing is written outside the destination's interior."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = os.path.join(root, "tools", "probe")
    if not os.path.exists(probe):   # normally built by __graft_entry__.build(); nvcc is on the GPU image too
        import sys
        sys.path.insert(0, root)
        import __graft_entry__
        __graft_entry__.build()
    assert os.path.exists(probe), "tools/probe is missing and could not be built (__graft_entry__.build())"
    r = subprocess.run([probe, "conv"] + shape.split(), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "-> PASS" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]


def test_utnet_batch_and_rect(utnet):
    sd = on.init_state_dict("UtNet", seed=0)
    torch.manual_seed(3)
    x = torch.rand(3, 3, 120, 136)
    with torch.no_grad():
        ref = on.utnet_forward(sd, x).numpy()
    y = utnet(x.to(dev())).cpu().numpy()
    check_pixels(y, ref, "UtNet batch 3, 120x136")
    # batch entries are independent
    y1 = utnet(x[1:2].to(dev())).cpu().numpy()
    assert np.abs(y1[0] - y[1]).max() <= 1e-6


def test_utnet_scaled_weights(utnet):
    """Weights rescaled so that the output has a lively range (sigma_out ~ 0.1-0.3), where an absolute
    2e-2 bound actually bites; also exercises re-packing after load_state_dict."""
    sd = on.init_state_dict("UtNet", seed=0)
    g = torch.Generator().manual_seed(11)
    sd2 = {k: v.clone() for k, v in sd.items()}
    for k in sd2:
        if k.endswith(".weight") and sd2[k].dim() == 4:
            sd2[k] = sd2[k] * 1.35
        if k.endswith(".bias"):
            sd2[k] = sd2[k] + 0.02 * torch.randn(sd2[k].shape, generator=g)
    m = nb.UtNet().to(dev()).eval()
    m.load_state_dict(sd2)
    torch.manual_seed(4)
    x = torch.rand(1, 3, 120, 120)
    with torch.no_grad():
        ref = on.utnet_forward(sd2, x).numpy()
    y = m(x.to(dev())).cpu().numpy()
    sigma = ref.std()
    err = np.abs(y - ref).max()
    print(f"scaled UtNet: sigma_out {sigma:.3f} max abs {err:.3e} rel {err / sigma:.3f}")
    assert err <= MAX_REL_SIGMA * sigma and np.isfinite(y).all()
    # in-place parameter update must trigger a re-pack
    with torch.no_grad():
        m.tconvs4[4].bias.add_(0.5)
    y2 = m(x.to(dev())).cpu().numpy()
    assert np.abs((y2 - y) - 0.5).max() <= 1e-4


def test_utnet_other_activations(golden_networks):
    for act in ("ELU", "Hardswish"):
        m = nb.UtNet(activation=act).to(dev()).eval()
        m.load_state_dict(on.init_state_dict("UtNet", seed=0, activation=act))
        torch.manual_seed(1)
        x = torch.rand(1, 3, 120, 120)
        y = m(x.to(dev())).cpu().numpy()[0]
        check_pixels(y, golden_networks[f"utnet_out_120_{act}"], f"UtNet {act}")


def test_utnet_illegal_crop_size(utnet):
    with pytest.raises(_capi.NindError, match="16a\\+56"):
        utnet(torch.rand(1, 3, 128, 128, device=dev()))


def test_unet_forward(unet, golden_networks):
    for cs in (64, 128):
        torch.manual_seed(1)
        x = torch.rand(1, 3, cs, cs)
        y = unet(x.to(dev())).cpu().numpy()[0]
        check_pixels(y, golden_networks[f"unet_out_{cs}"], f"UNet cs={cs} vs reference golden")
    sd = on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7)
    torch.manual_seed(2)
    x = torch.rand(2, 3, 96, 160)
    with torch.no_grad():
        ref = on.unet_forward(sd, x).numpy()
    check_pixels(unet(x.to(dev())).cpu().numpy(), ref, "UNet batch 2, 96x160")


# ------------------------------------------------------------------ whole image
def test_tiled_denoise_golden(utnet, golden_networks):
    N = golden_networks
    W, H, cs, ucs, ol = (int(v) for v in N["tiled_params"])
    img = torch.from_numpy(N["tiled_img"])
    out = nb.denoise_tiled(img.to(dev()), utnet, cs, ucs, ol, batch=4).cpu().numpy()
    check_pixels(out, N["tiled_out"], "tiled UtNet 300x260 vs reference loop")
    # PSNR difference against a clean target (north_star): here target = the noiseless input itself
    target = N["tiled_img"]
    d = abs(psnr(out, target) - psnr(N["tiled_out"], target))
    print(f"PSNR difference vs target: {d:.4f} dB")
    assert d <= PSNR_DIFF
    # host-buffer entry point gives the same image
    out_h = nb.denoise_tiled_host(img, utnet, cs, ucs, ol, batch=5).numpy()
    assert np.abs(out_h - out).max() <= 1e-6
    # sharded ranges sum to the whole
    n = nb.n_crops(W, H, cs, ucs, ol)
    acc = np.zeros_like(out)
    from nind_denoise_b200.tiler import _band
    for a, b in nb.shard_ranges(n, 4):
        if b > a:
            bd, y0, y1 = _band(utnet, img.to(dev()), cs, ucs, ol, a, b, 3)
            acc[:, y0:y1] += bd.cpu().numpy()
    assert np.abs(acc - out).max() <= 1e-6


def test_tiled_matches_gather_forward_stitch(utnet):
    """The fused tiled driver equals gather -> forward -> stitch done through the separate C-ABI ops
    (ties the bit-exact geometry ops to the fused gather+im2col path)."""
    W, H, cs, ucs, ol = 333, 271, 120, 96, 6
    torch.manual_seed(9)
    img = torch.rand(3, H, W, device=dev())
    fused = nb.denoise_tiled(img, utnet, cs, ucs, ol, batch=6)
    crops = nb.gather_crops(utnet, img, cs, ucs, ol)
    outs = torch.cat([utnet(crops[i:i + 6]) for i in range(0, crops.shape[0], 6)])
    band, _, _ = nb.stitch_crops(outs, H, W, cs, ucs, ol)
    assert float((band - fused).abs().max()) <= 1e-6


def test_kernel_launch_counter(utnet):
    n0 = _capi.lib().nind_kernel_launches()
    utnet(torch.rand(1, 3, 120, 120, device=dev()))
    torch.cuda.synchronize()
    # 1 gather + 22 conv launches; the four max-pools and the 1x1 head are fused into conv epilogues
    assert _capi.lib().nind_kernel_launches() - n0 == 23


def test_kernel_variants_agree():
    """Fused vs separate max-pool, pixel-pair vs plain 3x3 form, CTA-pair (cta_group::2) vs single-CTA tiles and
    flat vs 16x8 tiles compute the same function: identical bf16 pooling, and fp32 accumulation orders that
    differ only inside the MMA."""
    sd = on.init_state_dict("UtNet", seed=0)
    torch.manual_seed(6)
    x = torch.rand(2, 3, 120, 136, device=dev())
    outs = {}
    for name, opts in (("default", {}), ("unfused_pool", {"fuse_pool": 0}), ("no_pair", {"pair64": 0}),
                       ("cta1", {"cta_group": 1}),
                       ("cta2", {"cta_group": 2}), ("n128", {"n_tile_deep": 128}),
                       ("flat_off", {"flat": 0}), ("flat_all", {"flat": 1}),
                       ("wide_off", {"wide_store": 0}), ("wide_all", {"wide_store": 1}),
                       ("single_issuer", {"dual_issuer": 0})):
        m = nb.UtNet().to(dev()).eval()
        m.load_state_dict(sd)
        for k, v in opts.items():
            m.set_option(k, v)
        outs[name] = m(x).cpu().numpy()
    assert np.array_equal(outs["default"], outs["unfused_pool"])
    # pixel-pair mode (default for the C_out = 64 3x3 layers) walks K in a different order than the 3x3 form:
    # fp32 summation order differs inside the accumulator, which can flip bf16 roundings downstream
    d = np.abs(outs["no_pair"] - outs["default"]).max()
    print(f"pixel-pair vs 3x3 form of the C_out = 64 layers: max diff {d:.3e}")
    assert d <= 3e-4
    for k in ("cta1", "cta2", "n128"):
        assert np.abs(outs[k] - outs["default"]).max() <= 2e-5, k
    # flat (1-D) tiles only change which pixels share a tile, not any pixel's summation order
    assert np.array_equal(outs["flat_off"], outs["default"])
    assert np.array_equal(outs["flat_all"], outs["default"])
    # the staging layout of the TMA-store epilogue (64-byte halves / 128-byte rows) and the number of MMA issuer
    # warps do not touch the arithmetic at all
    for k in ("wide_off", "wide_all", "single_issuer"):
        assert np.array_equal(outs[k], outs["default"]), k


def test_cli_shim_roundtrip(tmp_path, golden_networks):
    """nind_denoise_b200.cli: the reference script's flags and file conventions
    (denoise_image.py:181-200; np_imgops.py:12-29; pt_helpers.py:22-40) around the GPU tiler."""
    import cv2
    from nind_denoise_b200 import cli

    N = golden_networks
    W, H, cs, ucs, ol = (int(v) for v in N["tiled_params"])
    img16 = (np.clip(N["tiled_img"], 0, 1) * 65535).round().astype(np.uint16)          # [3,H,W] RGB
    src = str(tmp_path / "in_s1.tif")
    cv2.imwrite(src, cv2.cvtColor(img16.transpose(1, 2, 0), cv2.COLOR_RGB2BGR))
    model_path = str(tmp_path / "generator_650.pt")
    sd = on.init_state_dict("UtNet", seed=0)
    torch.save(sd, model_path)
    dst = str(tmp_path / "out_s1_denoised.tiff")
    rc = cli.main(["--network", "UtNet", "--model_path", model_path, "--input", src, "--output", dst,
                   "--cs", str(cs), "--ucs", str(ucs), "--overlap", str(ol)])
    assert rc == 0
    got = cv2.cvtColor(cv2.imread(dst, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    assert got.dtype == np.float32 and got.shape == (3, H, W)
    x = img16.astype(np.float32) / 65535                                                # what the reader returns
    with torch.no_grad():
        ref = og.denoise_tiled(x, lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                               cs, ucs, ol)
    check_pixels(got, ref, "CLI shim (.tif in, float .tiff out)")
    # 16-bit output path clamps and quantises like pt_helpers.tensor_to_imgfile
    dst16 = str(tmp_path / "out.png")
    cli.main(["--network", "UtNet", "--model_path", model_path, "--input", src, "--output", dst16,
              "--cs", str(cs), "--ucs", str(ucs)])
    got16 = cv2.cvtColor(cv2.imread(dst16, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    assert got16.dtype == np.uint16
    assert np.abs(got16.astype(np.int64) - (np.clip(ref, 0, 1) * 65535).round().astype(np.int64)).max() <= 16


def test_unet_tiled_vs_oracle(unet):
    """BASELINE config 4 in miniature: UNet over a tiled image (cs multiple of 16, ucs = 0.75 cs)."""
    sd = on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7)
    W, H, cs, ucs, ol = 230, 170, 96, 72, 6
    img = np.random.default_rng(8).random((3, H, W), dtype=np.float32)
    out = nb.denoise_tiled(torch.from_numpy(img).to(dev()), unet, cs, ucs, ol, batch=5).cpu().numpy()
    with torch.no_grad():
        ref = og.denoise_tiled(img, lambda c: on.unet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                               cs, ucs, ol)
    check_pixels(out, ref, "tiled UNet 230x170")


def test_zero_overlap_and_large_crop(utnet):
    sd = on.init_state_dict("UtNet", seed=0)
    fwd = lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy()
    # overlap 0: no seam halving at all
    W, H, cs, ucs, ol = 250, 250, 120, 96, 0
    img = np.random.default_rng(9).random((3, H, W), dtype=np.float32)
    out = nb.denoise_tiled(torch.from_numpy(img).to(dev()), utnet, cs, ucs, ol).cpu().numpy()
    with torch.no_grad():
        ref = og.denoise_tiled(img, fwd, cs, ucs, ol)
    check_pixels(out, ref, "tiled UtNet overlap 0")
    # the largest benchmark crop size (nominal 1024 -> 1016), one crop through forward()
    torch.manual_seed(12)
    x = torch.rand(1, 3, 1016, 1016)
    with torch.no_grad():
        ref = on.utnet_forward(sd, x).numpy()
    check_pixels(utnet(x.to(dev())).cpu().numpy(), ref, "UtNet cs=1016")


def test_unet_reference_default_crop_and_odd_levels(unet):
    """The reference's own UNet tiling default is cs 440 (denoise_image.py:40): 440 -> 220 -> 110 -> 55 -> 27,
    i.e. an odd level where MaxPool2d floors and the upsampled tensor is zero-padded to the skip's size
    (ThirdPartyNets.py:114-118)."""
    sd = on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7)
    for shape in ((1, 3, 440, 440), (2, 3, 88, 104), (1, 3, 50, 70)):
        torch.manual_seed(13)
        x = torch.rand(*shape)
        with torch.no_grad():
            ref = on.unet_forward(sd, x).numpy()
        check_pixels(unet(x.to(dev())).cpu().numpy(), ref, f"UNet {shape}")


def test_crop_overlap_sweep(utnet):
    """BASELINE config 5's sweep in miniature: crop sizes x overlaps, tiled result vs the oracle loop."""
    sd = on.init_state_dict("UtNet", seed=0)
    fwd = lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy()
    img = np.random.default_rng(10).random((3, 301, 402), dtype=np.float32)
    dimg = torch.from_numpy(img).to(dev())
    for cs, ucs in ((120, 96), (136, 100), (248, 224)):
        for ol in (0, 6, 16, 32):
            out = nb.denoise_tiled(dimg, utnet, cs, ucs, ol).cpu().numpy()
            with torch.no_grad():
                ref = og.denoise_tiled(img, fwd, cs, ucs, ol)
            err = np.abs(out - ref).max()
            assert err <= MAX_ABS and err <= MAX_REL_SIGMA * ref.std(), (cs, ucs, ol, err)


def test_throughput_mode_matches_single_image_path(utnet):
    """nind_tiled_denoise_host_async + nind_host_sync: a stream of images (two device slots, three streams)
    gives exactly what the synchronous host entry gives image by image."""
    rng = np.random.default_rng(14)
    shapes = [(3, 260, 300), (3, 260, 300), (3, 333, 271), (3, 260, 300), (3, 200, 420)]
    imgs = [torch.from_numpy(rng.random(s, dtype=np.float32)).pin_memory() for s in shapes]
    cs, ucs, ol = 120, 96, 6
    outs = nb.denoise_images_host(imgs, utnet, cs, ucs, ol, batch=5)
    for im, out in zip(imgs, outs):
        ref = nb.denoise_tiled_host(im, utnet, cs, ucs, ol, batch=4)
        assert out.shape == im.shape
        assert float((out - ref).abs().max()) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3, 5])
def test_host_range_pipeline_composes_to_the_whole_image(utnet, world):
    """nind_tiled_denoise_host_range + nind_host_join_rows: the per-rank share of the multi-GPU host entry, run here
    rank after rank on one GPU (last rank first, so that every hand-over exists when its owner needs it) with the
    ownership and seam plan of denoise_tiled_distributed_host, gives the single-GPU image; rows [o0, hi) arrive in
    the host image straight from the pipeline, and the rows an earlier rank owns are final on a side stream that
    only waited for the step that completes them."""
    from nind_denoise_b200.tiler import host_range
    rng = np.random.default_rng(21)
    H, W, cs, ucs, ol = 610, 455, 120, 96, 6
    img = torch.from_numpy(rng.random((3, H, W), dtype=np.float32)).pin_memory()
    ref = nb.denoise_tiled_host(img, utnet, cs, ucs, ol, batch=7)
    ranges = nb.shard_ranges(nb.n_crops(W, H, cs, ucs, ol), world)
    ext = nb.band_extents(W, H, cs, ucs, ol, ranges)
    own = nb.owned_rows_up(ext, H)
    out = torch.full((3, H, W), float("nan")).pin_memory()
    side = torch.cuda.Stream()
    handed = {}   # (sender, owner) -> rows
    for r in reversed(range(world)):
        cb, ce = ranges[r]
        if ce <= cb:
            continue
        o0, o1 = own[r]
        sends, recvs = nb.seam_plan(ext, own, r)
        hi = max(o0, min([o1] + [a for _, a, _ in recvs]))
        full = host_range(utnet, img, out, cs, ucs, ol, 6, cb, ce, o0, hi)
        if sends:
            _capi.check(_capi.lib().nind_host_join_rows(utnet.native_handle(), max(b for _, _, b in sends),
                                                        C.c_void_p(side.cuda_stream)))
            with torch.cuda.stream(side):
                for q, a, b in sends:
                    handed[(r, q)] = full[:, a:b].clone()
            side.synchronize()
        for q, a, b in recvs:   # rank order
            full[:, a:b] += handed[(q, r)]
        out[:, hi:o1] = full[:, hi:o1].cpu()
        _capi.check(_capi.lib().nind_host_sync(utnet.native_handle()))
        torch.cuda.synchronize()
        assert not torch.isnan(out[:, o0:o1]).any()
    assert not torch.isnan(out).any()
    assert float((out - ref).abs().max()) <= 1e-6


@pytest.mark.gpu
def test_dir_cli_matches_single_image_path(tmp_path):
    """nind_denoise_b200.dir_cli (denoise_dir.py:76-129 without the per-image subprocess): a directory
    streamed through the async host entry gives, file by file, what the single-image CLI writes; the lowest-ISO
    file is the baseline (dataset_torch_3.get_baseline_fpath) and the averaged mse / ssim / msssim losses land in
    testres.json next to the model (json_saver.py)."""
    import json

    import cv2
    from nind_denoise_b200 import cli, dir_cli, scoring

    rng = np.random.default_rng(31)
    noisy = tmp_path / "set_200_176"
    noisy.mkdir()
    shapes = {"NIND_x_ISO800.tif": (180, 170), "NIND_x_ISO3200.png": (180, 170), "NIND_x_ISO6400.tif": (180, 170)}
    for name, (h, w) in shapes.items():
        cv2.imwrite(str(noisy / name), (rng.random((h, w, 3)) * 65535).astype(np.uint16))
    cv2.imwrite(str(noisy / "NIND_x_ISO200.tif"), (rng.random((180, 170, 3)) * 65535).astype(np.uint16))
    mdir = tmp_path / "2021_model"
    mdir.mkdir()
    model_path = str(mdir / "generator_7.pt")
    torch.save(on.init_state_dict("UtNet", seed=0), model_path)
    rc = dir_cli.main(["--noisy_dir", str(noisy), "--result_dir", str(tmp_path / "out"), "--network", "UtNet",
                       "--model_path", model_path, "--cs", "120", "--ucs", "96"])
    assert rc == 0
    out_dir = tmp_path / "out" / "2021_model"          # denoise_dir.py:59: result_dir / <model directory name>
    assert sorted(os.listdir(out_dir)) == sorted(shapes)  # the ISO200 baseline is not denoised
    clean = torch.from_numpy(cli.img_path_to_np_flt(str(noisy / "NIND_x_ISO200.tif")))
    per_img = []
    for name in shapes:
        single = str(tmp_path / ("single_" + name))
        cli.main(["--network", "UtNet", "--model_path", model_path, "--input", str(noisy / name), "--output", single,
                  "--cs", "120", "--ucs", "96", "--exif_method", "noexif"])
        a = cv2.imread(str(out_dir / name), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(single, cv2.IMREAD_UNCHANGED)
        assert a.dtype == np.uint16 and a.shape == b.shape, name
        assert np.abs(a.astype(np.int64) - b.astype(np.int64)).max() <= 1, name   # same kernels, same batches
        per_img.append(scoring.get_losses(clean, torch.from_numpy(cli.img_path_to_np_flt(str(out_dir / name)))))
    res = json.load(open(mdir / "testres.json"))
    assert set(res["7"]) == {"test_mse", "test_ssim", "test_msssim"} and res["best_epoch"]["test_mse"] == 7
    for k in ("mse", "ssim", "msssim"):  # the scores are those of the files on disk, as pt_helpers.get_losses computes
        assert abs(res["7"]["test_" + k] - np.mean([d[k] for d in per_img])) <= 1e-5, k


@pytest.mark.gpu
def test_cli_jpeg_in_jpeg_out_and_whole_image(tmp_path):
    """The remaining file conventions of denoise_image.py through the CLI shim: an 8-bit JPEG in (np_imgops.py:19-22:
    /255) and a .jpg out (pt_helpers.py:36-40: clip(0,1), 8 bit) on the tiled path, and --whole_image --pad with a
    16-bit .png out, both against the oracle pushed through the same cv2 encoder."""
    import cv2
    from nind_denoise_b200 import cli

    sd = {k: v.clone() for k, v in on.init_state_dict("UtNet", seed=0).items()}
    sd["tconvs4.4.bias"] = sd["tconvs4.4.bias"] + 0.5     # output inside the displayable range
    model_path = str(tmp_path / "generator_1.pt")
    torch.save(sd, model_path)
    rng = np.random.default_rng(61)
    yy, xx = np.meshgrid(np.linspace(0, 1, 230), np.linspace(0, 1, 260), indexing="ij")
    smooth = np.stack([0.5 + 0.4 * np.sin(5 * xx), 0.5 + 0.4 * np.cos(4 * yy), 0.5 + 0.3 * np.sin(3 * xx + 2 * yy)], -1)
    src = str(tmp_path / "in.jpg")
    cv2.imwrite(src, np.clip((smooth + rng.normal(0, 0.03, smooth.shape)) * 255, 0, 255).astype(np.uint8))
    img = cli.img_path_to_np_flt(src)                      # what the reference reads from that file: RGB CHW / 255
    assert img.dtype == np.float32 and img.shape == (3, 230, 260) and img.max() <= 1.0

    def fwd(c):
        return on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy()

    # tiled, .jpg out
    out_jpg = str(tmp_path / "out.jpg")
    assert cli.main(["--network", "UtNet", "--model_path", model_path, "--input", src, "--output", out_jpg,
                     "--cs", "120", "--ucs", "96", "--exif_method", "noexif"]) == 0
    with torch.no_grad():
        ref = og.denoise_tiled(img, fwd, 120, 96, 6)
    exp8 = (np.clip(ref, 0, 1) * 255 + 0.5).astype(np.uint8)                       # pt_helpers.py:36-38
    exp_path = str(tmp_path / "expected.jpg")
    cv2.imwrite(exp_path, cv2.cvtColor(np.ascontiguousarray(exp8.transpose(1, 2, 0)), cv2.COLOR_RGB2BGR))
    got, exp = cv2.imread(out_jpg), cv2.imread(exp_path)
    assert got.shape == exp.shape == (230, 260, 3) and got.dtype == np.uint8
    # a pixel whose value sits on an 8-bit rounding boundary may differ by one level before encoding
    # (and the JPEG encoder spreads such a flip over its 8x8 block)
    d = np.abs(got.astype(np.int32) - exp.astype(np.int32))
    assert d.max() <= 4 and d.mean() <= 0.25, (int(d.max()), float(d.mean()))

    # whole image, 16-bit .png out
    small = str(tmp_path / "small.png")
    cv2.imwrite(small, (rng.random((104, 136, 3)) * 65535).astype(np.uint16))
    out_png = str(tmp_path / "whole.png")
    assert cli.main(["--network", "UtNet", "--model_path", model_path, "--input", small, "--output", out_png,
                     "--whole_image", "--pad", "8", "--cs", "120", "--ucs", "96", "--exif_method", "noexif"]) == 0
    im2 = cli.img_path_to_np_flt(small)
    with torch.no_grad():
        y = fwd(og.whole_image_input(im2, 8))[:, 8:-8, 8:-8]
    exp16 = (np.clip(y, 0, 1) * 65535).round().astype(np.int64)
    got16 = cv2.cvtColor(cv2.imread(out_png, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).transpose(2, 0, 1).astype(np.int64)
    assert got16.shape == exp16.shape
    assert np.abs(got16 - exp16).max() <= 0.25 * float(y.std()) * 65535 + 1


@pytest.mark.gpu
def test_dir_cli_two_gpus_match_one(tmp_path):
    """dir_cli --gpus 2 (one replica process per GPU fed from one file queue — what replaces denoise_dir.py:76-103's
    per-image subprocesses on a multi-GPU box) writes the same files and the same testres.json scores as --gpus 1.
    Skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json

    import cv2
    from nind_denoise_b200 import dir_cli

    rng = np.random.default_rng(33)
    noisy = tmp_path / "set_a"
    noisy.mkdir()
    names = [f"NIND_a_ISO{iso}.tif" for iso in (400, 800, 1600, 3200, 6400)]
    for name in names + ["NIND_a_ISO100.tif"]:
        cv2.imwrite(str(noisy / name), (rng.random((150, 190, 3)) * 65535).astype(np.uint16))
    outs = {}
    for gpus in (1, 2):
        mdir = tmp_path / f"model_g{gpus}"
        mdir.mkdir()
        model_path = str(mdir / "generator_3.pt")
        torch.save(on.init_state_dict("UtNet", seed=0), model_path)
        rc = dir_cli.main(["--noisy_dir", str(noisy), "--result_dir", str(tmp_path / f"out{gpus}"), "--network", "UtNet",
                           "--model_path", model_path, "--cs", "120", "--ucs", "96", "--gpus", str(gpus)])
        assert rc == 0
        d = tmp_path / f"out{gpus}" / f"model_g{gpus}"
        assert sorted(os.listdir(d)) == sorted(names)
        outs[gpus] = ({n: cv2.imread(str(d / n), cv2.IMREAD_UNCHANGED) for n in names},
                      json.load(open(mdir / "testres.json"))["3"])
    for n in names:
        assert np.array_equal(outs[1][0][n], outs[2][0][n]), n
    for k, v in outs[1][1].items():
        assert abs(v - outs[2][1][k]) <= 1e-7, k


@pytest.mark.gpu
def test_whole_image_mode(utnet):
    """--whole_image (denoise_image.py:91-97,110-128,255-256): one forward over the mirror-padded image."""
    sd = on.init_state_dict("UtNet", seed=0)
    rng = np.random.default_rng(8)
    for (h, w, pad) in ((104, 136, 8), (120, 120, 0)):
        img = rng.random((3, h, w), dtype=np.float32)
        got = nb.denoise_whole_image(torch.from_numpy(img).to(dev()), utnet, pad).cpu().numpy()
        x = torch.from_numpy(og.whole_image_input(img, pad)).unsqueeze(0)
        with torch.no_grad():
            ref = on.utnet_forward(sd, x)[0].numpy()[:, pad:h + pad, pad:w + pad]
        assert got.shape == (3, h, w)
        check_pixels(got, ref, f"whole image {h}x{w} pad {pad}")
    with pytest.raises(Exception):   # padded size 100+16 is not a legal UtNet size
        nb.denoise_whole_image(torch.zeros(3, 100, 100, device=dev()), utnet, 8)


# ------------------------------------------------------------------ headline configurations (round 2)
def exclusive_region_errors(img, out, sd, network, cs, ucs, ol, crops):
    """For each sampled crop: the stitched pixels it owns exclusively (useful area minus the seam bands shared with
    neighbours) against the fp32 oracle forward of that crop.  Returns (max abs error, sigma of the oracle values)."""
    H, W = img.shape[1], img.shape[2]
    g = og.crop_grid(W, H, cs, ucs, ol)
    fwd = on.utnet_forward if network == "UtNet" else on.unet_forward
    err, refs = 0.0, []
    with torch.no_grad():
        for i in crops:
            e = og.crop_entry(g, i)
            y = fwd(sd, torch.from_numpy(og.gather_crop(img, g, i)).unsqueeze(0))[0].numpy()
            xlo, ylo, xhi, yhi = e["usefuldim"]
            ax, ay = e["usefulstart"]
            h, w = yhi - ylo, xhi - xlo
            l, t = (ol if ax != 0 else 0), (ol if ay != 0 else 0)
            r = ol if (ax + ucs < W and ol) else 0
            b = ol if (ay + ucs < H and ol) else 0
            ref = y[:, ylo + t:yhi - b, xlo + l:xhi - r]
            got = out[:, ay + t:ay + h - b, ax + l:ax + w - r]
            assert got.shape == ref.shape and ref.size > 0
            err = max(err, float(np.abs(got.astype(np.float64) - ref).max()))
            refs.append(ref.ravel())
    return err, float(np.concatenate(refs).std())


def corner_edge_interior(nx, ny):
    return [0, nx - 1, (ny - 1) * nx, nx * ny - 1, nx // 2, (ny // 2) * nx + nx // 2]


@pytest.mark.parametrize("cs", [248, 504])
def test_headline_24mp_image_against_oracle(utnet, cs):
    """BASELINE configs[1]: the full 6000x4000 image at the benchmarked plan (default batch), six sampled crops
    (corners, an edge, the interior) against the oracle, and equality with a batch-4 run."""
    W, H, ucs, ol = 6000, 4000, cs - 24, 6
    sd = on.init_state_dict("UtNet", seed=0)
    img = torch.rand((3, H, W), generator=torch.Generator().manual_seed(1))
    dimg = img.to(dev())
    out = nb.denoise_tiled(dimg, utnet, cs, ucs, ol)          # default batch = bench.py's plan
    g = og.crop_grid(W, H, cs, ucs, ol)
    crops = corner_edge_interior(g.nx, g.ny) if cs == 248 else [0, g.size - 1, (g.ny // 2) * g.nx + g.nx // 2]
    err, sigma = exclusive_region_errors(img.numpy(), out.cpu().numpy(), sd, "UtNet", cs, ucs, ol, crops)
    print(f"24 MP cs {cs}: max abs {err:.3e}, sigma_out {sigma:.4f}, err/sigma {err / sigma:.3f}")
    assert err <= MAX_ABS and err <= MAX_REL_SIGMA * sigma
    out4 = nb.denoise_tiled(dimg, utnet, cs, ucs, ol, batch=4)
    assert float((out4 - out).abs().max()) <= 1e-6
    # the host-buffer entry (what bench.py's e2e leg calls) gives the same image
    out_h = nb.denoise_tiled_host(img.pin_memory(), utnet, cs, ucs, ol)
    assert float((out_h - out.cpu()).abs().max()) <= 1e-6


def test_unet_cs512_multi_row_grid(unet):
    """BASELINE configs[3] in shape: UNet at cs 512 / ucs 384 over a grid with several rows and columns."""
    sd = on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7)
    W, H, cs, ucs, ol = 1000, 900, 512, 384, 6
    img = np.random.default_rng(41).random((3, H, W), dtype=np.float32)
    g = og.crop_grid(W, H, cs, ucs, ol)
    assert g.nx >= 2 and g.ny >= 2
    out = nb.denoise_tiled(torch.from_numpy(img).to(dev()), unet, cs, ucs, ol).cpu().numpy()
    err, sigma = exclusive_region_errors(img, out, sd, "UNet", cs, ucs, ol, [0, g.nx - 1, g.size - 1, g.nx + 1])
    print(f"UNet cs 512 {g.nx}x{g.ny} grid: max abs {err:.3e}, err/sigma {err / sigma:.3f}")
    assert err <= MAX_ABS and err <= MAX_REL_SIGMA * sigma
    # whole image against the oracle loop where seams are summed too (smaller image keeps the oracle quick)
    W2, H2 = 700, 560
    img2 = np.random.default_rng(42).random((3, H2, W2), dtype=np.float32)
    out2 = nb.denoise_tiled(torch.from_numpy(img2).to(dev()), unet, cs, ucs, ol).cpu().numpy()
    with torch.no_grad():
        ref2 = og.denoise_tiled(img2, lambda c: on.unet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                                cs, ucs, ol)
    check_pixels(out2, ref2, "tiled UNet cs 512, 700x560")


def test_utnet_cs504_tiled_vs_oracle(utnet):
    """The reference's own default tiling (cs 504 / ucs 480, denoise_image.py:41) on a two-by-two grid."""
    sd = on.init_state_dict("UtNet", seed=0)
    W, H, cs, ucs, ol = 900, 700, 504, 480, 6
    img = np.random.default_rng(43).random((3, H, W), dtype=np.float32)
    out = nb.denoise_tiled(torch.from_numpy(img).to(dev()), utnet, cs, ucs, ol).cpu().numpy()
    with torch.no_grad():
        ref = og.denoise_tiled(img, lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                               cs, ucs, ol)
    check_pixels(out, ref, "tiled UtNet cs 504, 900x700")


def test_psnr_difference_against_a_clean_target(utnet):
    """north_star: PSNR difference <= 0.05 dB versus the fp32 reference, measured against a CLEAN target: a smooth
    synthetic image plus sigma = 0.05 Gaussian noise, clipped to 0..1.  The weights are the default init plus an
    output-bias shift, so that the result lands in the image's range and PSNR-to-target is a meaningful number."""
    sd = {k: v.clone() for k, v in on.init_state_dict("UtNet", seed=0).items()}
    sd["tconvs4.4.bias"] = sd["tconvs4.4.bias"] + 0.6
    m = nb.UtNet().to(dev()).eval()
    m.load_state_dict(sd)
    W, H, cs, ucs, ol = 520, 430, 248, 224, 6
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    clean = np.stack([0.5 + 0.35 * np.sin(6.0 * xx + 1.0) * np.cos(4.0 * yy),
                      0.5 + 0.30 * np.sin(5.0 * yy + 0.5 * xx),
                      0.45 + 0.25 * np.cos(7.0 * xx * yy)]).astype(np.float32)
    noisy = np.clip(clean + np.random.default_rng(44).normal(0, 0.05, clean.shape), 0, 1).astype(np.float32)
    out = nb.denoise_tiled(torch.from_numpy(noisy).to(dev()), m, cs, ucs, ol).cpu().numpy()
    with torch.no_grad():
        ref = og.denoise_tiled(noisy, lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                               cs, ucs, ol)
    check_pixels(out, ref, "tiled UtNet on a noisy smooth image")
    d = abs(psnr(out, clean) - psnr(ref, clean))
    print(f"PSNR vs clean target: b200 {psnr(out, clean):.3f} dB, fp32 oracle {psnr(ref, clean):.3f} dB, diff {d:.4f} dB")
    assert d <= PSNR_DIFF


# ------------------------------------------------------------------ round-2 entry points
def test_denoise_batch_clamps_like_the_reference(utnet):
    """Generator.denoise_batch = model(x).clip(0, 1) (nn_common.py:198-199); the clamp is fused into the head."""
    sd = {k: v.clone() for k, v in on.init_state_dict("UtNet", seed=0).items()}
    for k in sd:
        if k.endswith(".weight") and sd[k].dim() == 4:
            sd[k] = sd[k] * 1.35          # lively output range (sigma_out ~ 0.1-0.3)
    torch.manual_seed(5)
    x = torch.rand(2, 3, 120, 136)
    with torch.no_grad():
        ref0 = on.utnet_forward(sd, x)
    # centre the output on 0 through the head's bias, so that the clamp bites on about half of the pixels
    shift = ref0.mean(dim=(0, 2, 3))
    sd["tconvs4.4.bias"] = sd["tconvs4.4.bias"] - shift
    ref = ref0 - shift.view(1, 3, 1, 1)
    m = nb.UtNet().to(dev()).eval()
    m.load_state_dict(sd)
    y = m.denoise_batch(x.to(dev())).cpu()
    assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    assert float((ref < 0).float().mean()) > 0.05 and float((ref > 0).float().mean()) > 0.05
    err = float((y - ref.clip(0, 1)).abs().max())
    assert err <= MAX_REL_SIGMA * float(ref.std())
    assert torch.equal(y, m(x.to(dev())).cpu().clip(0, 1))   # identical arithmetic, clamp aside
    # UNet: sigmoid output is inside (0, 1) already; the flag must be harmless
    un = nb.UNet().to(dev()).eval()
    un.load_state_dict(on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7))
    xu = torch.rand(1, 3, 64, 80, device=dev())
    assert torch.equal(un.denoise_batch(xu), un(xu))


def test_file_format_kernels_bit_exact():
    """nind_image_to_chw_f32 / nind_chw_f32_to_image against the reference expressions
    (np_imgops.py:19-28, pt_helpers.py:24-32)."""
    from nind_denoise_b200 import cli
    rng = np.random.default_rng(51)
    h, w = 37, 53
    for dtype, scale in ((np.uint8, 255), (np.uint16, 65535)):
        bgr = rng.integers(0, scale + 1, (h, w, 3)).astype(dtype)
        bgr[0, 0] = (0, scale, 1)
        t = torch.from_numpy(bgr.view(np.int16) if dtype == np.uint16 else bgr).to(dev())
        t = t.view(torch.uint16) if dtype == np.uint16 else t
        got = cli.image_to_chw(t, bgr=True).cpu().numpy()
        ref = bgr[:, :, ::-1].transpose(2, 0, 1).astype(np.single) / scale
        assert np.array_equal(got, ref), dtype
        got_rgb = cli.image_to_chw(t, bgr=False).cpu().numpy()
        assert np.array_equal(got_rgb, bgr.transpose(2, 0, 1).astype(np.single) / scale)
    f = rng.random((h, w, 3), dtype=np.float32) * 3 - 1     # stage-1 TIFFs keep highlights > 1 (denoise.py:417)
    assert np.array_equal(cli.image_to_chw(torch.from_numpy(f).to(dev())).cpu().numpy(), f[:, :, ::-1].transpose(2, 0, 1))
    x = torch.from_numpy((rng.random((3, h, w), dtype=np.float32) * 1.4 - 0.2))
    x[0, 0, :4] = torch.tensor([0.5 / 65535, 1.5 / 65535, 2.5 / 65535, 1.0])     # round-half-even cases
    q16 = cli.chw_to_image(x.to(dev()), torch.uint16).view(torch.int16).cpu().numpy().view(np.uint16)
    ref16 = (x.clip(0, 1) * 65535).round().numpy().astype(np.uint16).transpose(1, 2, 0)[:, :, ::-1]
    assert np.array_equal(q16, ref16)
    q8 = cli.chw_to_image(x.to(dev()), torch.uint8).cpu().numpy()
    ref8 = (x.clip(0, 1) * 255).add(0.5).clamp(0, 255).byte().numpy().transpose(1, 2, 0)[:, :, ::-1]
    assert np.array_equal(q8, ref8)
    qf = cli.chw_to_image(x.to(dev()), torch.float32).cpu().numpy()
    assert np.array_equal(qf, x.numpy().transpose(1, 2, 0)[:, :, ::-1])


def test_cli_model_directory_and_float_tiff(tmp_path, golden_networks):
    """--model_path given as a directory (Model.complete_path, nn_common.py:75-114: best epoch from trainres.json)
    and a float32 stage-1 TIFF in, float32 .tiff out (the production route of denoise.py:397-436)."""
    import json

    import cv2
    from nind_denoise_b200 import cli

    N = golden_networks
    W, H, cs, ucs, ol = (int(v) for v in N["tiled_params"])
    src = str(tmp_path / "in_s1.tif")
    cv2.imwrite(src, cv2.cvtColor(N["tiled_img"].transpose(1, 2, 0), cv2.COLOR_RGB2BGR))      # 32-bit float TIFF
    mdir = tmp_path / "models" / "run1"
    mdir.mkdir(parents=True)
    sd = on.init_state_dict("UtNet", seed=0)
    torch.save(sd, str(mdir / "generator_12.pt"))
    torch.save({k: v * 0 for k, v in sd.items()}, str(mdir / "generator_30.pt"))            # a worse, later epoch
    json.dump({"best_epoch": {"validation_loss": 12}}, open(mdir / "trainres.json", "w"))
    dst = str(tmp_path / "out.tiff")
    rc = cli.main(["--network", "UtNet", "--model_path", "run1", "--models_dpath", str(tmp_path / "models"),
                   "--input", src, "--output", dst, "--cs", str(cs), "--ucs", str(ucs), "--exif_method", "noexif"])
    assert rc == 0
    got = cv2.cvtColor(cv2.imread(dst, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    # models_dpath/name drops the keyword (nn_common.py:111): the highest-numbered file wins there ...
    assert np.abs(got).max() <= 1e-6
    # ... while the directory itself honours trainres.json's best epoch
    rc = cli.main(["--network", "UtNet", "--model_path", str(mdir), "--input", src, "--output", dst, "--cs", str(cs),
                   "--ucs", str(ucs), "--exif_method", "noexif"])
    got = cv2.cvtColor(cv2.imread(dst, cv2.IMREAD_UNCHANGED), cv2.COLOR_BGR2RGB).transpose(2, 0, 1)
    check_pixels(got, N["tiled_out"], "CLI, float TIFF in / out, model directory")


def test_entry_points_on_different_streams_do_not_race(utnet):
    """The entry points of one handle share scratch memory (plan arenas, crop outputs, origin table); calls
    enqueued back to back on different streams — and on the library's own host-pipeline streams — are ordered by
    the handle's last-use event."""
    rng = np.random.default_rng(61)
    cs, ucs, ol = 120, 96, 6
    a = torch.from_numpy(rng.random((3, 500, 620), dtype=np.float32)).to(dev())
    b = torch.from_numpy(rng.random((3, 500, 620), dtype=np.float32)).to(dev())
    bh = b.cpu().pin_memory()
    ref_a = nb.denoise_tiled(a, utnet, cs, ucs, ol, batch=20).clone()
    ref_b = nb.denoise_tiled(b, utnet, cs, ucs, ol, batch=20).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            out_a = nb.denoise_tiled(a, utnet, cs, ucs, ol, batch=20)
        with torch.cuda.stream(s2):
            out_b = nb.denoise_tiled(b, utnet, cs, ucs, ol, batch=20)
        out_bh = nb.denoise_tiled_host(bh, utnet, cs, ucs, ol, batch=20)      # library streams, synchronises itself
        with torch.cuda.stream(s1):
            out_a2 = nb.denoise_tiled(a, utnet, cs, ucs, ol, batch=20)
        torch.cuda.synchronize()
        assert torch.equal(out_a, ref_a) and torch.equal(out_b, ref_b) and torch.equal(out_a2, ref_a)
        assert float((out_bh - ref_b.cpu()).abs().max()) <= 1e-6


def test_module_moved_to_another_gpu():
    """.to(another device) after the first forward: the native handle is rebuilt on the new device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sd = on.init_state_dict("UtNet", seed=0)
    m = nb.UtNet().to("cuda:0").eval()
    m.load_state_dict(sd)
    x = torch.rand(1, 3, 120, 120)
    y0 = m(x.to("cuda:0")).cpu()
    m.to("cuda:1")
    y1 = m(x.to("cuda:1")).cpu()
    assert torch.equal(y0, y1)
    m2 = nb.UtNet().to("cuda:0").eval()          # a second handle on another device of the same process
    m2.load_state_dict(sd)
    assert torch.equal(m2(x.to("cuda:0")).cpu(), y0) and torch.equal(m(x.to("cuda:1")).cpu(), y0)


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    d = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=d)
    m = nb.UtNet().to(d).eval()
    m.load_state_dict(on.init_state_dict("UtNet", seed=0))
    W, H, cs, ucs, ol = 1210, 990, 120, 96, 6
    img = torch.rand((3, H, W), generator=torch.Generator().manual_seed(3)).pin_memory()
    out = nb.denoise_tiled_distributed(img.to(d), m, cs, ucs, ol)
    for _ in range(3):   # the peer-memory gather alternates between two output images
        out_p = nb.denoise_tiled_distributed(img.to(d), m, cs, ucs, ol, mode="peer")
    sh = nb.SharedHostImage((3, H, W))
    for _ in range(2):
        out_h = nb.denoise_tiled_distributed_host(img, m, cs, ucs, ol, out=sh)
    if rank == 0:
        single = nb.denoise_tiled(img.to(d), m, cs, ucs, ol)
        q.put((max(float((out - single).abs().max()), float((out_p - single).abs().max())),
               float((out_h - single.cpu()).abs().max())))
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


def test_nccl_sharded_equals_single_gpu():
    """BASELINE configs[2] under pytest: crops sharded over 2 NCCL ranks (device-resident gather to rank 0 and the
    shared-host-image path) equal the single-GPU image.  Skipped on a one-GPU box (bench.py's `parity` block
    makes the same comparison on every multi-GPU bench run)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    d_dev, d_host = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert d_dev <= 1e-6 and d_host <= 1e-6


@pytest.mark.parametrize("world", [2, 3, 8])
def test_step_api_composes_to_the_whole_image(utnet, world):
    """nind_plan_steps / nind_tiled_denoise_step / nind_add_rows — what mode="peer" of the multi-GPU gather is made
    of — run rank after rank on one GPU: owned rows copied into the output, partial sums for rows later ranks own
    added in rank order, equals the single-GPU image."""
    rng = np.random.default_rng(71)
    H, W, cs, ucs, ol = 730, 655, 120, 96, 6
    img = torch.from_numpy(rng.random((3, H, W), dtype=np.float32)).to(dev())
    ref = nb.denoise_tiled(img, utnet, cs, ucs, ol, batch=9)
    ranges = nb.shard_ranges(nb.n_crops(W, H, cs, ucs, ol), world)
    ext = nb.band_extents(W, H, cs, ucs, ol, ranges)
    own = nb.owned_rows(ext, H)
    out = torch.full((3, H, W), float("nan"), device=dev())
    seams = []
    for r, (cb, ce) in enumerate(ranges):
        if ce <= cb:
            seams.append(None)
            continue
        full = torch.full((3, H, W), float("nan"), device=dev())
        steps = nb.plan_steps(utnet, W, H, cs, ucs, ol, cb, ce, 7)
        assert steps[0][0] == cb and steps[-1][1] == ce and all(a[1] == b[0] for a, b in zip(steps, steps[1:]))
        done = ext[r][0]
        for (a, b) in steps:
            r0, r1 = nb.tiled_step(utnet, img, full, cs, ucs, ol, cb, ce, a, b)
            assert r0 == done and r1 >= r0
            done = r1
        assert done == ext[r][1]
        o0, o1 = own[r]
        out[:, o0:o1] = full[:, o0:o1]
        seams.append(full[:, o1:ext[r][1]].clone())
    for r, s in enumerate(seams):
        if s is not None and s.shape[1] > 0:
            nb.add_rows(out[:, own[r][1]:ext[r][1]], s)
    assert not torch.isnan(out).any()
    assert float((out - ref).abs().max()) <= 1e-6
