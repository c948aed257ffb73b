"""Kernel (c) parity: fused CFG + x0 + PLMS/DDIM + concat (C-ABI) vs the scheduler oracle, fp32, 2e-6."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = dict(rtol=2e-6, atol=2e-6)


def px(t):  # NCHW -> pixel-major [B*h*w, 4]
    return t.permute(0, 2, 3, 1).reshape(-1, 4).contiguous()


def unpx(t, B, h, w):
    return t.reshape(B, h, w, 4).permute(0, 3, 1, 2).contiguous()


def _run_dual(sched_cls, oracle_cls, steps, B=2, h=8, w=8, g=7.5, phi=0.0, eta=0.0):
    from gm_diffusion_b200 import schedulers as S
    from oracle import schedulers_oracle as O
    gen = torch.Generator().manual_seed(1)
    lat = torch.randn(B, 4, h, w, generator=gen)
    so, go = oracle_cls(), None
    so.set_timesteps(steps)
    go = copy.deepcopy(so)
    sp = sched_cls(); sp.set_timesteps(steps)
    gp = S.clone_scheduler(sp)
    assert torch.equal(sp.timesteps, so.timesteps)
    n_px = B * h * w
    sdr, gm = S.BranchState(n_px, "cuda"), S.BranchState(n_px, "cuda")
    sdr.x.copy_(px(lat)); gm.x.copy_(px(lat))
    x_o, gm_o = lat.clone(), lat.clone()
    unet_in = torch.full((n_px, 8), 7.0, dtype=torch.bfloat16, device="cuda")
    gm_in = torch.full((n_px, 8), 7.0, dtype=torch.bfloat16, device="cuda")
    ws = torch.zeros(B * 4, device="cuda")
    for i, t in enumerate(sp.timesteps.tolist()):
        eu, ec, eg = (torch.randn(B, 4, h, w, generator=gen) for _ in range(3))
        nz = torch.randn(B, 4, h, w, generator=gen), torch.randn(B, 4, h, w, generator=gen)
        # oracle (dual_unet.py:1063-1093)
        e = eu + g * (ec - eu)
        if phi > 0:
            e = O.rescale_noise_cfg(e, ec, phi)
        a = so.alphas_cumprod[t]
        x0 = (x_o - (1 - a).sqrt() * e) / a.sqrt()
        ancestral = oracle_cls.__name__ == "DDPMOracle"
        kw = dict(eta=eta, variance_noise=nz[0]) if eta > 0 else (dict(variance_noise=nz[0]) if ancestral else {})
        x_o_next = so.step(e, t, x_o, **kw)[0]
        gm_in_o = torch.cat([x0, gm_o], 1)
        kw = dict(eta=eta, variance_noise=nz[1]) if eta > 0 else (dict(variance_noise=nz[1]) if ancestral else {})
        gm_o = go.step(eg, t, gm_o, **kw)[0]
        x_o = x_o_next
        # product
        plan = sp.plan_step(t, eta)
        if plan.needs_noise or ancestral:
            sdr.noise = px(nz[0]).cuda(); gm.noise = px(nz[1]).cuda()
        S.fused_step(plan, sdr, px(ec).cuda(), px(eu).cuda(), guidance_scale=g, guidance_rescale=phi, px_per_sample=h * w,
                     x0_coeffs=sp.x0_coeffs(t), unet_in_next=unet_in, concat_out=gm_in, concat_tail=gm.x, rescale_ws=ws)
        S.fused_step(gp.plan_step(t, eta), gm, px(eg).cuda(), x0_coeffs=gp.x0_coeffs(t))
        torch.testing.assert_close(unpx(sdr.x.cpu(), B, h, w), x_o, **TOL, msg=lambda m: f"sdr step {i} t={t}: {m}")
        torch.testing.assert_close(unpx(gm.x.cpu(), B, h, w), gm_o, **TOL, msg=lambda m: f"gm step {i} t={t}: {m}")
        gi = gm_in.float().cpu().reshape(B, h, w, 8).permute(0, 3, 1, 2)
        torch.testing.assert_close(gi, gm_in_o.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
        ui = unet_in.float().cpu().reshape(B, h, w, 8)
        torch.testing.assert_close(ui[..., :4].permute(0, 3, 1, 2), x_o.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
        assert float(ui[..., 4:].abs().max()) == 0.0
    return x_o, gm_o


@pytest.mark.parametrize("steps", [4, 10, 50])
def test_plms_dual(steps):
    from gm_diffusion_b200.schedulers import PNDMScheduler
    from oracle.schedulers_oracle import PNDMOracle
    _run_dual(PNDMScheduler, PNDMOracle, steps)


def test_plms_guidance_rescale():
    from gm_diffusion_b200.schedulers import PNDMScheduler
    from oracle.schedulers_oracle import PNDMOracle
    _run_dual(PNDMScheduler, PNDMOracle, 6, phi=0.7)


@pytest.mark.parametrize("eta", [0.0, 0.7])
def test_ddim_dual(eta):
    from gm_diffusion_b200.schedulers import DDIMScheduler
    from oracle.schedulers_oracle import DDIMOracle
    _run_dual(DDIMScheduler, DDIMOracle, 10, eta=eta)


@pytest.mark.parametrize("steps", [5, 50])
def test_ddpm_dual(steps):
    """The scheduler the reference CLIs really pass (generate_hdr.py:162-176): ancestral noise, SDR draw then GM draw."""
    from gm_diffusion_b200.schedulers import DDPMScheduler
    from oracle.schedulers_oracle import DDPMOracle
    _run_dual(DDPMScheduler, DDPMOracle, steps)


@pytest.mark.parametrize("steps", [4, 25, 50])
def test_dpmsolver_dual(steps):
    """DPM-Solver++(2M): the scheduler formal_improved.py:195 swaps in; the history ring holds x0 predictions."""
    from gm_diffusion_b200.schedulers import DPMSolverMultistepScheduler
    from oracle.schedulers_oracle import DPMSolverOracle
    _run_dual(DPMSolverMultistepScheduler, DPMSolverOracle, steps)


def test_latent_layout_roundtrip_and_pack():
    import ctypes as C
    from gm_diffusion_b200 import _lib as L
    x = torch.randn(3, 4, 5, 7, device="cuda")
    p = torch.empty(3 * 35, 4, device="cuda")
    y = torch.empty_like(x)
    L.check(L.lib().gmd_latents_nchw_to_px(x.data_ptr(), p.data_ptr(), 3, 35, L.current_stream()))
    assert torch.equal(p.reshape(3, 5, 7, 4).permute(0, 3, 1, 2), x)
    L.check(L.lib().gmd_latents_px_to_nchw(p.data_ptr(), y.data_ptr(), 3, 35, L.current_stream()))
    assert torch.equal(x, y)
    out = torch.empty(105, 16, dtype=torch.bfloat16, device="cuda")
    L.check(L.lib().gmd_pack_unet_input(p.data_ptr(), p.data_ptr(), out.data_ptr(), 105, 16, L.current_stream()))
    assert torch.equal(out[:, :4], p.to(torch.bfloat16)) and torch.equal(out[:, 4:8], p.to(torch.bfloat16)) and float(out[:, 8:].abs().max()) == 0


def test_empty_and_errors():
    from gm_diffusion_b200 import schedulers as S
    st = S.BranchState(0, "cuda")
    sp = S.PNDMScheduler(); sp.set_timesteps(4)
    S.fused_step(sp.plan_step(751), st, torch.empty(0, 4, device="cuda"))
    st = S.BranchState(16, "cuda")
    with pytest.raises(ValueError):
        S.fused_step(sp.plan_step(501), st, torch.zeros(16, 4, device="cuda"), unet_in_next=torch.zeros(16, 4, dtype=torch.bfloat16, device="cuda"))
