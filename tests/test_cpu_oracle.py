"""CPU suite, part 1: the oracle against the reference's own outputs (golden fixtures generated from the real
tone_mapping.py) and the analytic pins of SURVEY.md §8c."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import schedulers_oracle as SO
from oracle import tone_mapping_oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir / "tm_reference.npz")


@pytest.mark.parametrize("qmax", [9, 49, 99])
def test_tm_oracle_matches_reference_fixtures(gold, qmax):
    sdr, gm = torch.from_numpy(gold["sdr"]), torch.from_numpy(gold["gm"])
    hdr = O.apply_gm_to_sdr(gm, sdr, float(qmax), 1 / 64)
    assert torch.equal(hdr, torch.from_numpy(gold[f"hdr_q{qmax}"]))
    assert torch.equal(O.linear_scale_tmo(hdr, qmax), torch.from_numpy(gold[f"linear_q{qmax}"]))
    assert torch.equal(O.hard_clip_tmo(hdr, qmax), torch.from_numpy(gold[f"hardclip_q{qmax}"]))
    assert torch.equal(O.fix_mulog_tmo(hdr, qmax), torch.from_numpy(gold[f"mulog_q{qmax}"]))
    assert torch.equal(O.tmo_cuda(hdr), torch.from_numpy(gold[f"tmo_cuda_q{qmax}"]))
    torch.testing.assert_close(O.gamut_compress(O.fix_mulog_tmo(hdr, qmax)), torch.from_numpy(gold[f"mulog_gamut_q{qmax}"]), rtol=1e-6, atol=1e-7)


def test_tm_oracle_spot_values(golden_dir):
    g = np.load(golden_dir / "tm_spot.npz")
    hdr = O.apply_gm_to_sdr(torch.from_numpy(g["gm"]), torch.from_numpy(g["sdr"]), 99.0)
    assert torch.equal(hdr, torch.from_numpy(g["hdr"]))
    assert abs(float(hdr.max()) - 85.567) < 1e-3
    assert abs(float(O.fix_mulog_tmo(hdr, 99).mean()) - 0.55577) < 1e-4
    assert abs(float(O.gamut_compress(O.fix_mulog_tmo(hdr, 99)).mean()) - 0.53887) < 1e-4


@pytest.mark.skipif(not os.path.exists(O.REFERENCE_TM), reason="reference checkout not present (GPU box)")
def test_tm_oracle_against_live_reference():
    ref = O.load_reference_tm()
    g = torch.Generator().manual_seed(42)
    sdr, gm = torch.rand(2, 3, 17, 13, generator=g) * 1.4 - 0.2, torch.rand(2, 3, 17, 13, generator=g)
    for q in (9.0, 49.0, 99.0):
        h = ref.apply_gm_to_sdr(gm, sdr, q)
        assert torch.equal(O.apply_gm_to_sdr(gm, sdr, q), h)
        assert torch.equal(O.fix_mulog_tmo(h, q), ref.fix_mulog_tmo(h, q))
        assert torch.equal(O.linear_scale_tmo(h, q), ref.linear_scale_tmo(h, q))
        assert torch.equal(O.hard_clip_tmo(h, q), ref.hard_clip_tmo(h, q))
        assert torch.equal(O.gamut_compress(O.fix_mulog_tmo(h, q)), ref.gamut_compress(ref.fix_mulog_tmo(h, q)))
    assert torch.equal(O.tmo_cuda(sdr * 20), ref.tmo_cuda(sdr * 20))
    with pytest.raises(ValueError):
        O.tmo_cuda(torch.tensor([float("nan")]))


def test_tm_identities_and_numpy_twin():
    s = torch.linspace(0, 1, 101)
    torch.testing.assert_close(O.apply_gm_to_sdr(torch.zeros_like(s), s, 99.0), s ** 2.2, rtol=1e-6, atol=1e-7)
    assert float(O.apply_gm_to_sdr(torch.ones(1), torch.ones(1), 99.0)) == 100.0
    assert float(O.fix_mulog_tmo(torch.tensor([100.0]), 99)) == 1.0 and float(O.fix_mulog_tmo(torch.tensor([0.0]), 99)) == 0.0
    M = torch.tensor(O.BT2020_TO_709)
    torch.testing.assert_close(M.sum(1), torch.ones(3), rtol=0, atol=2e-6)       # neutral grey is preserved
    rng = np.random.default_rng(1)
    a, b = rng.random((5, 7, 3), dtype=np.float32), rng.random((5, 7, 3), dtype=np.float32)
    np.testing.assert_allclose(O.apply_gm_to_sdr_numpy(b, a, 99.0), O.apply_gm_to_sdr(torch.from_numpy(b), torch.from_numpy(a), 99.0).numpy(), rtol=2e-6, atol=1e-6)


# ---- scheduler pins -------------------------------------------------------------------------
def test_pndm_timestep_literals():
    s = SO.PNDMOracle()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-2:] == [21, 1]
    s.set_timesteps(4)
    assert s.timesteps.tolist() == [751, 501, 501, 251, 1]
    d = SO.DDIMOracle(); d.set_timesteps(4)
    assert d.timesteps.tolist() == [751, 501, 251, 1]
    assert abs(float(s.alphas_cumprod[0]) - 0.999150) < 1e-6 and abs(float(s.alphas_cumprod[981]) - 0.005775) < 1e-5


def test_plms_step_equals_ddim_step_for_same_eps():
    p, d = SO.PNDMOracle(), SO.DDIMOracle()
    p.set_timesteps(10); d.set_timesteps(10)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64), torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    p.alphas_cumprod = p.alphas_cumprod.double(); d.alphas_cumprod = d.alphas_cumprod.double()
    a = p._get_prev_sample(x, 501, 401, e)
    b = d.step(e, 501, x)[0]
    torch.testing.assert_close(a, b, rtol=1e-10, atol=1e-10)


def test_constant_eps_trajectory_closed_form():
    """With eps constant every PLMS combination returns eps, so x_t = sqrt(a_t) x0 + sqrt(1-a_t) eps is preserved."""
    for cls in (SO.PNDMOracle, SO.DDIMOracle):
        s = cls()
        s.alphas_cumprod = s.alphas_cumprod.double()
        s.final_alpha_cumprod = s.alphas_cumprod[0]
        s.set_timesteps(50)
        g = torch.Generator().manual_seed(3)
        x0, e = torch.randn(1, 4, 4, 4, generator=g, dtype=torch.float64), torch.randn(1, 4, 4, 4, generator=g, dtype=torch.float64)
        a = s.alphas_cumprod[int(s.timesteps[0])]
        x = a.sqrt() * x0 + (1 - a).sqrt() * e
        for t in s.timesteps:
            x = s.step(e, t, x)[0]
        af = s.alphas_cumprod[0]
        torch.testing.assert_close(x, af.sqrt() * x0 + (1 - af).sqrt() * e, rtol=1e-8, atol=1e-8)


def test_unet_oracle_parameter_counts_and_shapes():
    from oracle.unet_oracle import UNet2DConditionOracle, count_params, widen_conv_in
    from oracle.vae_oracle import VaeDecoderOracle
    with torch.device("meta"):
        assert count_params(UNet2DConditionOracle(4)) == 859_520_964      # the public SD1.5 figure
        assert count_params(UNet2DConditionOracle(8)) == 859_532_484      # + 4*320*9 (train_gm_unet.py:658-677)
    torch.manual_seed(0)
    small = UNet2DConditionOracle(4, block_out_channels=(32, 64, 64, 64), cross_attention_dim=48).eval()
    x, ctx = torch.randn(2, 4, 16, 16), torch.randn(2, 77, 48)
    with torch.no_grad():
        y = small(x, 501, encoder_hidden_states=ctx)
        y64 = small.double()(x.double(), 501, encoder_hidden_states=ctx.double())
    assert y.shape == x.shape
    assert float((y.double() - y64).norm() / y64.norm()) < 1e-5           # fp32 vs fp64 self-consistency
    wide = widen_conv_in(small.float())
    with torch.no_grad():
        y8 = wide(torch.cat([x, x], 1), 501, encoder_hidden_states=ctx)
    torch.testing.assert_close(y8, y, rtol=1e-4, atol=1e-5)               # tiled/halved conv_in: f([x,x]) == f(x)
    v = VaeDecoderOracle(ch=(32, 64, 64, 64)).eval()
    with torch.no_grad():
        assert v.decode(torch.randn(1, 4, 4, 4)).shape == (1, 3, 32, 32)
    # full AutoencoderKL: the public SD VAE figures (83 653 863 parameters = encoder 34 163 592 + decoder 49 490 179 + 72 + 20)
    from oracle.vae_oracle import VaeOracle
    with torch.device("meta"):
        full = VaeOracle()
        assert count_params(full.encoder) == 34_163_592 and count_params(full.decoder) == 49_490_179
        assert count_params(full) == 83_653_863
    torch.manual_seed(0)
    e = VaeOracle(ch=(32, 64, 64, 64)).eval()
    with torch.no_grad():
        d = e.encode(torch.randn(1, 3, 32, 48))
        assert d.mean.shape == (1, 4, 4, 6) and torch.equal(d.mode(), d.mean)
        z = d.sample(generator=torch.Generator().manual_seed(1))
        n = torch.randn(1, 4, 4, 6, generator=torch.Generator().manual_seed(1))
        torch.testing.assert_close(z, d.mean + torch.exp(0.5 * d.logvar) * n)
    from gm_diffusion_b200.random_init import sd_vae_state_dict
    sd = sd_vae_state_dict(0, device="cpu", block_out_channels=(32, 64, 64, 64))
    want = {k: tuple(t.shape) for k, t in e.state_dict().items()}
    assert {k: tuple(t.shape) for k, t in sd.items()} == want  # the bench's random-init VAE has diffusers' key names and shapes


def test_oracle_loops_run_and_keep_quirks():
    from oracle import pipeline_oracle as PO
    from oracle.unet_oracle import UNet2DConditionOracle, widen_conv_in
    torch.manual_seed(0)
    u4 = UNet2DConditionOracle(4, block_out_channels=(32, 64, 64, 64), cross_attention_dim=48).eval()
    u8 = widen_conv_in(u4)
    pe, ne = torch.randn(1, 77, 48), torch.randn(1, 77, 48)
    lat = torch.randn(1, 4, 8, 8)
    trace = []
    a, b = PO.dual_unet_loop(u4, u8, SO.PNDMOracle(), pe, ne, lat, num_inference_steps=4, trace=trace)
    assert [s["t"] for s in trace] == [751, 501, 501, 251, 1] and a.shape == b.shape == lat.shape
    assert not torch.equal(a, b)
    c = PO.single_gm_loop(u8, SO.PNDMOracle(), 0.18215 * torch.randn(1, 4, 8, 8), pe, ne, lat, num_inference_steps=4)
    assert c.shape == lat.shape and torch.isfinite(c).all()


# ---- host tail: Radiance RGBE as cv2.imwrite quantises it (generate_hdr.py:27-30) ----------------------------------------
def test_rgbe_oracle_matches_real_cv2_files(golden_dir):
    """oracle/rgbe_oracle.py is pinned to bytes parsed out of files the real cv2.imwrite wrote (oracle/make_golden.py)."""
    from oracle import rgbe_oracle as RO
    g = np.load(golden_dir / "rgbe_cv2.npz")
    for name in ("wide", "narrow"):
        want = g[f"{name}_rgbe"]
        got = RO.save_hdr_pixels(g[f"{name}_hdr"], 99)
        assert np.array_equal(got, want), f"{name}: {int((got != want).any(-1).sum())} pixels differ"
        assert np.array_equal(RO.rgbe2float(want), g[f"{name}_decoded"])
    # edge pixels: below the 1e-32 cut-off -> all-zero RGBE; exact powers of two sit at mantissa 128
    px = RO.float2rgbe(np.array([[0, 0, 0], [9e-33, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125]], np.float32))
    assert px.tolist() == [[0, 0, 0, 0], [0, 0, 0, 0], [128, 128, 128, 129], [128, 64, 32, 128]]


def test_radiance_container_roundtrip_and_cv2_read(golden_dir, tmp_path):
    """gm_diffusion_b200.hdr_io.pack_radiance (host code, no GPU): parse(pack(x)) == x for RLE-able and flat widths, and the real
    cv2.imread decodes our file to exactly what it decodes the reference-written file to."""
    from gm_diffusion_b200.hdr_io import pack_radiance
    from oracle import rgbe_oracle as RO
    rng = np.random.default_rng(3)
    for H, W in ((3, 8), (2, 127), (2, 128), (3, 129), (1, 300), (4, 5), (2, 1)):
        x = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        x[0, 0] = (2, 2, 0, W & 255)  # a first pixel that looks like an RLE marker must survive
        assert np.array_equal(RO.parse_radiance(pack_radiance(x)), x), (H, W)
    with pytest.raises(ValueError):
        pack_radiance(np.zeros((4, 4, 3), np.uint8))
    cv2 = pytest.importorskip("cv2")
    g = np.load(golden_dir / "rgbe_cv2.npz")
    for name in ("wide", "narrow"):
        path = tmp_path / f"{name}.hdr"
        path.write_bytes(pack_radiance(g[f"{name}_rgbe"]))
        back = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)[:, :, ::-1]
        assert np.array_equal(back, g[f"{name}_decoded"]), name


def test_tm_oracle_gradients_match_reference_autograd(golden_dir):
    """The stage-1 training chain is differentiated with autograd in the reference (train_vqgan_lora.py:1134-1141); the oracle's
    autograd gradients must equal the fixtures taken from the REAL tone_mapping.py (oracle/make_golden.py:grad_golden)."""
    g = np.load(golden_dir / "tm_grads.npz")
    w = torch.from_numpy(g["w"])
    gm = torch.from_numpy(g["gm"]).requires_grad_(True)
    sdr = torch.from_numpy(g["sdr"]).requires_grad_(True)
    y = O.gamut_compress(O.fix_mulog_tmo(O.apply_gm_to_sdr(gm, sdr, qmax=49), 49))
    (y * w).sum().backward()
    np.testing.assert_allclose(y.detach().numpy(), g["chain_out"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gm.grad.numpy(), g["chain_g_gm"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(sdr.grad.numpy(), g["chain_g_sdr"], rtol=1e-5, atol=1e-7)


def test_exposure_oracle_matches_reference_fixture(golden_dir):
    """oracle/augment_oracle.py vs outputs of the REAL augmentations.py for the same seeds (same RNG consumption, same arithmetic)."""
    import random
    from oracle import augment_oracle as AO
    g = np.load(golden_dir / "exposure_reference.npz")
    x = torch.from_numpy(g["x"])
    random.seed(3); torch.manual_seed(3)
    for i in range(4):
        y, meta = AO.random_exposure_adjust(x)
        assert [meta["exposure"], meta["n"], meta["sigma"]] == g[f"meta{i}"].tolist()
        assert np.array_equal(y.numpy(), g[f"y{i}"])
    assert np.array_equal(AO.apply_inv_sigmoid_curve(x, 0.7, 0.55).numpy(), g["curve"])
    assert np.array_equal(AO.discretize_to_uint16(x * 1.2 - 0.1).numpy(), g["quant"])
    assert np.array_equal(AO.hdr_to_ldr(x * 3, 0.5).numpy(), g["ldr"])
