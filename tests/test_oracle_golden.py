"""CPU: the oracle restatement against the golden vectors produced by the reference itself
(oracle/make_golden.py).  These pin the oracle; the GPU parity tests then compare against the oracle."""
import os

import numpy as np
import torch

from oracle import geometry as og
from oracle import nets as on

N_GEOM, N_TABLE = 7, 10


def fake_model_np(c):
    x = torch.from_numpy(c).unsqueeze(0)
    ramp = torch.linspace(0.5, 1.5, x.shape[-1]).view(1, 1, 1, -1) * torch.linspace(1.25, 0.75, x.shape[-2]).view(1, 1, -1, 1)
    return (x * ramp + 0.125)[0].numpy()


def test_crop_tables_and_crops(golden_geometry):
    G = golden_geometry
    for gi in range(N_GEOM):
        W, H, cs, ucs, ol = (int(v) for v in G[f"g{gi}_params"])
        g = og.crop_grid(W, H, cs, ucs, ol)
        table = G[f"g{gi}_table"]
        assert g.size == table.shape[0]
        assert np.array_equal(og.crop_table(g)[:, 2:], table)
        img = G[f"g{gi}_img"]
        for k, i in enumerate(G[f"g{gi}_crop_idx"]):
            assert np.array_equal(og.gather_crop(img, g, int(i)), G[f"g{gi}_crops"][k])


def test_stitch_bit_exact(golden_geometry):
    G = golden_geometry
    for gi in range(N_GEOM):
        W, H, cs, ucs, ol = (int(v) for v in G[f"g{gi}_params"])
        out = og.denoise_tiled(G[f"g{gi}_img"], fake_model_np, cs, ucs, ol)
        assert np.array_equal(out, G[f"g{gi}_stitched"]), gi


def test_baseline_grid_sizes(golden_geometry):
    G = golden_geometry
    expect = {(6000, 4000, 504, 480, 6): 117, (6000, 4000, 248, 224, 6): 532, (8256, 5504, 512, 384, 6): 330,
              (6000, 4000, 120, 96, 6): 3015, (6000, 4000, 1016, 992, 6): 35}
    for ti in range(N_TABLE):
        W, H, cs, ucs, ol = (int(v) for v in G[f"t{ti}_params"])
        g = og.crop_grid(W, H, cs, ucs, ol)
        assert g.size == int(G[f"t{ti}_n"][0])
        if (W, H, cs, ucs, ol) in expect:
            assert g.size == expect[(W, H, cs, ucs, ol)]
        t = og.crop_table(g)
        assert np.array_equal(t[[0, -1], 2:], G[f"t{ti}_first_last"])


def test_seam_weights_sum_to_one():
    # with cs-ucs even the halved seams sum to exactly 1 everywhere
    for (W, H, cs, ucs, ol) in [(101, 83, 40, 28, 4), (300, 260, 120, 96, 6), (64, 64, 40, 28, 0)]:
        out = og.denoise_tiled(np.ones((3, H, W), np.float32), lambda c: np.ones_like(c), cs, ucs, ol)
        assert np.array_equal(out, np.ones_like(out))


def test_utnet_forward_golden(golden_networks):
    N = golden_networks
    sd = on.init_state_dict("UtNet", seed=0)
    chk = np.array([float(v.double().abs().sum()) for v in sd.values()])
    assert np.allclose(chk, N["utnet_sd_checksum"], rtol=0, atol=0)
    assert len(sd) == 64 and sum(v.numel() for v in sd.values()) == 31031893
    with torch.no_grad():
        for cs in (120, 248):
            torch.manual_seed(1)
            x = torch.rand(1, 3, cs, cs)
            y = on.utnet_forward(sd, x)[0].numpy()
            assert np.abs(y - N[f"utnet_out_{cs}"]).max() <= 1e-6


def test_utnet_other_activations_golden(golden_networks):
    with torch.no_grad():
        for act in ("ELU", "Hardswish"):
            sd = on.init_state_dict("UtNet", seed=0, activation=act)
            assert len(sd) == 46
            torch.manual_seed(1)
            x = torch.rand(1, 3, 120, 120)
            y = on.utnet_forward(sd, x, activation=act)[0].numpy()
            assert np.abs(y - golden_networks[f"utnet_out_120_{act}"]).max() <= 1e-6


def test_unet_forward_golden(golden_networks):
    sd = on.randomize_bn_(on.init_state_dict("UNet", seed=0), seed=7)
    assert len(sd) == 136
    with torch.no_grad():
        for cs in (64, 128):
            torch.manual_seed(1)
            x = torch.rand(1, 3, cs, cs)
            y = on.unet_forward(sd, x)[0].numpy()
            assert np.abs(y - golden_networks[f"unet_out_{cs}"]).max() <= 1e-6


def test_tiled_utnet_golden(golden_networks):
    N = golden_networks
    W, H, cs, ucs, ol = (int(v) for v in N["tiled_params"])
    sd = on.init_state_dict("UtNet", seed=0)
    with torch.no_grad():
        out = og.denoise_tiled(N["tiled_img"], lambda c: on.utnet_forward(sd, torch.from_numpy(c).unsqueeze(0))[0].numpy(),
                               cs, ucs, ol)
    assert np.abs(out - N["tiled_out"]).max() <= 1e-6


def test_flops_closed_form():
    for cs, g in [(120, 14.22357248), (248, 73.974119936), (504, 338.00254208), (1016, 1444.168695296)]:
        assert abs(on.utnet_flops(cs) / 1e9 - g) < 1e-6


def test_whole_image_input_matches_reference_golden():
    """oracle.geometry.whole_image_input == the reference's OneImageDS whole-image branch (vectors generated
    by oracle/make_golden.py from the reference itself), and the host layer's torch version == the oracle on
    non-square images too."""
    import nind_denoise_b200 as nb
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "whole_image.npz"))
    for k in range(3):
        img, pad = G[f"w{k}_img"], int(G[f"w{k}_pad"][0])
        assert np.array_equal(og.whole_image_input(img, pad), G[f"w{k}_input"])
        assert np.array_equal(nb.pad_whole_image(torch.from_numpy(img), pad).numpy(), G[f"w{k}_input"])
    rng = np.random.default_rng(2)
    for (h, w, pad) in ((20, 31, 5), (17, 9, 9), (12, 40, 0)):
        img = rng.random((3, h, w), dtype=np.float32)
        assert np.array_equal(nb.pad_whole_image(torch.from_numpy(img), pad).numpy(), og.whole_image_input(img, pad))
