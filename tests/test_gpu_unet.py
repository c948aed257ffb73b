"""UNet / VAE executors (sm_100a kernels through the C-ABI) vs the fp32 oracle restatements on identical random-init
weights and seeded inputs.  north_star tolerance: per-step UNet eps within 1e-2 relative L2 (bf16 compute)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="module")
def oracle_unets():
    from oracle.unet_oracle import UNet2DConditionOracle, widen_conv_in
    torch.manual_seed(0)
    u4 = UNet2DConditionOracle(4).eval()
    u8 = widen_conv_in(u4).eval()
    return u4.cuda(), u8.cuda()


@pytest.fixture(scope="module")
def b200_unets(oracle_unets):
    from gm_diffusion_b200 import B200UNet
    return tuple(B200UNet.from_module(u) for u in oracle_unets)


@pytest.mark.parametrize("which,B,hw,t", [(0, 2, 32, 981), (0, 1, 64, 501), (1, 2, 32, 21), (1, 3, 16, 741)])
def test_unet_forward_parity(oracle_unets, b200_unets, which, B, hw, t):
    ref, mine = oracle_unets[which], b200_unets[which]
    g = torch.Generator().manual_seed(100 + hw + t)
    cin = 4 if which == 0 else 8
    x = torch.randn(B, cin, hw, hw, generator=g).cuda()
    ctx = torch.randn(B, 77, 768, generator=g).cuda()
    with torch.no_grad():
        want = ref(x, t, encoder_hidden_states=ctx)
    got = mine.forward_nchw(x, t, ctx)
    assert got.shape == want.shape and got.dtype == torch.float32
    r = rel_l2(got, want)
    assert r < 1e-2, f"eps rel-L2 {r:.3e} (in={cin}, B={B}, {hw}x{hw}, t={t})"


def test_unet_batch_independence(b200_unets):
    mine = b200_unets[0]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 4, 32, 32, generator=g).cuda()
    ctx = torch.randn(3, 77, 768, generator=g).cuda()
    full = mine.forward_nchw(x, 500, ctx)
    one = mine.forward_nchw(x[1:2], 500, ctx[1:2])
    assert torch.equal(full[1:2], one), f"a sample's eps must not depend on the batch it is in (rel {rel_l2(full[1:2], one):.2e})"


def test_cfg_shared_prefix_is_exact(b200_unets):
    """cfg_shared (one copy of the latents for both CFG halves, prefix computed once) == the duplicated batch, bit for bit."""
    mine = b200_unets[0]
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 32, 32, 8, generator=g).to(torch.bfloat16).cuda()
    x[..., 4:] = 0
    ctx = torch.randn(4, 77, 768, generator=g).cuda()
    kv = mine.project_context(ctx)
    tb = mine.timestep_table([321])
    full = mine.forward(torch.cat([x, x]), tb, kv)
    shared = mine.forward(x, tb, kv, cfg_shared=True)
    assert shared.shape == full.shape and torch.equal(shared, full)


def test_vae_decoder_parity():
    from gm_diffusion_b200 import B200VaeDecoder
    from oracle.vae_oracle import VaeDecoderOracle
    torch.manual_seed(1)
    ref = VaeDecoderOracle().eval().cuda()
    mine = B200VaeDecoder.from_module(ref)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, 4, 16, 16, generator=g).cuda()
    with torch.no_grad():
        want = ref.decode(z)
    got = mine.decode(z)
    assert got.shape == want.shape
    r = rel_l2(got, want)
    # Error budget: ~40 bf16 conv / GroupNorm layers at up to 512 channels on random-init weights (no trained contraction of errors).
    # The yardstick is the SAME decoder executed in bf16 by torch (cuDNN), i.e. the reference's own reduced-precision path: measured
    # 2.7-3.6e-2 over three seeds against 1.8-2.4e-2 here (fp32 GroupNorm inputs, fp32 residual adds in the epilogues).  The gate is
    # therefore relative — at least 20 % better than torch-bf16 — plus the absolute 3e-2; the HDR PSNR gate covers the end-to-end effect.
    with torch.no_grad():
        bf = ref.to(torch.bfloat16).decode(z.to(torch.bfloat16))
    r_bf16 = rel_l2(bf, want)
    assert r < 3e-2 and r < 0.8 * r_bf16, f"VAE decode rel-L2 {r:.3e} (torch bf16 of the same decoder: {r_bf16:.3e})"


def test_vae_encoder_parity_and_latent_dist():
    """`pipeline.vae.encode(sdr_image).latent_dist.sample()` (generate_hdr.py:207-209) on the B200 kernels vs the fp32 oracle."""
    from gm_diffusion_b200 import B200Vae
    from gm_diffusion_b200 import _lib as L
    from oracle.vae_oracle import VaeOracle
    torch.manual_seed(2)
    ref = VaeOracle().eval().cuda()
    mine = B200Vae.from_module(ref)
    g = torch.Generator().manual_seed(5)
    img = (torch.rand(2, 3, 128, 192, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        want = ref.moments(img)                                   # [B,8,h,w]
    L.lib().gmd_reset_launch_count()
    dist = mine.encode(img).latent_dist
    assert L.lib().gmd_launch_count() > 50                        # the encoder ran on the library's kernels
    assert dist.mean.shape == (2, 4, 16, 24) and dist.parameters.dtype == torch.float32
    r_mean = rel_l2(dist.mean, want[:, :4])
    r_all = rel_l2(dist.parameters.permute(0, 3, 1, 2), want)
    assert r_mean < 1.5e-2 and r_all < 1.5e-2, f"encoder moments rel-L2 mean {r_mean:.3e} all {r_all:.3e}"   # measured 1.0e-2
    # sample / mode: exact DiagonalGaussianDistribution arithmetic on the product's own moments
    noise = torch.randn(2, 4, 16, 24, generator=g).cuda()
    got = dist.sample(noise=noise)
    lv = dist.logvar.clamp(-30.0, 20.0)
    torch.testing.assert_close(got, dist.mean + torch.exp(0.5 * lv) * noise, rtol=2e-6, atol=2e-6)
    assert torch.equal(dist.mode(), dist.mean.contiguous())
    a = dist.sample(generator=torch.Generator().manual_seed(9))   # CPU generator: CPU draw moved to the device, like randn_tensor
    b = dist.sample(noise=torch.randn(2, 4, 16, 24, generator=torch.Generator().manual_seed(9)).cuda())
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        mine.encode(torch.zeros(1, 3, 100, 128, device="cuda"))
    # encode -> decode keeps the image geometry (SDR -> latent -> image)
    assert mine.decode(dist.mode()).shape == (2, 3, 128, 192)
