"""CPU suite, part 2: host logic of the product (no compute calls): the C-ABI library loads and exports every symbol the
header declares, struct layouts match, scheduler plans equal the oracle's coefficients, argument checking mirrors the
reference, sharding arithmetic, and a world_size-2 gloo run of the gather path."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from gm_diffusion_b200 import build
    build.build()
    from gm_diffusion_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    L = lib.lib()
    names = lib.declared_symbols()
    assert len(names) >= 17 and "gmd_attn_fwd" in names and "gmd_hdr_reconstruct" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/gmd_b200.h but not exported"
    assert L.gmd_version() == 201
    assert abs(L.gmd_decode_ordered(0x3F800000) - 1.0) == 0.0 and L.gmd_decode_ordered(-2139095041) == float("-inf")


def test_struct_layouts_match_header(lib):
    src = r'''
    #include <stdio.h>
    #include "gmd_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(gmd_hdr_params), sizeof(gmd_sched_params), sizeof(gmd_gemm_params),
                            sizeof(gmd_conv_params), sizeof(gmd_attn_params)); return 0; }'''
    import tempfile
    tmp = Path(tempfile.mkdtemp(prefix="gmd_sizes_"))
    (tmp / "sizes.c").write_text(src)
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(tmp / "sizes.c"), "-o", str(tmp / "sizes")], check=True)
    out = subprocess.run([str(tmp / "sizes")], check=True, capture_output=True, text=True).stdout.split()
    want = [C.sizeof(lib.HdrParams), C.sizeof(lib.SchedParams), C.sizeof(lib.GemmParams), C.sizeof(lib.ConvParams), C.sizeof(lib.AttnParams)]
    assert [int(x) for x in out] == want


def test_argument_errors_without_gpu(lib):
    L = lib.lib()
    p = lib.HdrParams()
    assert L.gmd_hdr_reconstruct(C.byref(p), None) == -1 and b"null input" in L.gmd_last_error()
    with pytest.raises(ValueError):
        lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        lib.check(-3, "x")
    with pytest.raises(RuntimeError):
        lib.check(-2, "x")
    a = lib.AttnParams()
    assert L.gmd_attn_fwd(C.byref(a), None) == -1
    g = lib.GemmParams()
    assert L.gmd_gemm_fwd(C.byref(g), None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_product_fails_loudly_without_cuda():
    import gm_diffusion_b200 as G
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        G.apply_gm_to_sdr(torch.rand(4), torch.rand(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        G.StableDiffusionGMPipeline(None, None, None, None, G.PNDMScheduler())


def test_no_oracle_import_in_product():
    for f in (ROOT / "gm_diffusion_b200").rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"
    assert "gm_diffusion_b200" not in sys.modules or "oracle" not in getattr(sys.modules["gm_diffusion_b200"], "__dict__", {})
    # measurement scripts under profiles/ do not import it either: the GPU-side comparators receive the oracle's torch modules from
    # `bench.py --comparators` (bench.py is the one script besides tests/ and smoke() that may execute oracle/)
    for f in (ROOT / "profiles").glob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f"{f} imports the oracle"


def test_bench_cli_declares_the_contract_flags():
    """bench.py keeps the driver's flags (--gpus/--steps/--warmup/--impl) next to the comparator switch."""
    txt = (ROOT / "bench.py").read_text()
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--comparators", "--no-cpu-baseline"):
        assert f'"{flag}"' in txt, flag


@pytest.mark.parametrize("steps", [4, 10, 50])
def test_pndm_plan_matches_oracle_coefficients(steps):
    from gm_diffusion_b200.schedulers import PNDMScheduler
    from oracle.schedulers_oracle import PNDMOracle
    p, o = PNDMScheduler(), PNDMOracle()
    p.set_timesteps(steps); o.set_timesteps(steps)
    assert torch.equal(p.timesteps, o.timesteps) and torch.equal(p.alphas_cumprod, o.alphas_cumprod)
    kinds = []
    x = torch.ones(1)
    for t in p.timesteps.tolist():
        plan = p.plan_step(t)
        kinds.append(plan.plms_kind)
        # feeding eps = 1, sample = 1 exposes the (sample, eps) coefficients of the oracle's linear update
        e = torch.ones(1)
        o_out = o.step(e, t, x)[0]
        src = 1.0
        mine = plan.c_sample * src - plan.c_num * 1.0 / plan.c_denom
        assert abs(float(o_out) - mine) < 1e-6, (t, float(o_out), mine)
    assert kinds[:5] == [0, 1, 2, 3, 4][: len(kinds[:5])] and all(k == 4 for k in kinds[4:])


def test_ddim_plan_matches_oracle():
    from gm_diffusion_b200.schedulers import DDIMScheduler
    from oracle.schedulers_oracle import DDIMOracle
    p, o = DDIMScheduler(), DDIMOracle()
    p.set_timesteps(10); o.set_timesteps(10)
    assert torch.equal(p.timesteps, o.timesteps)
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(8, generator=g) for _ in range(3))
    for eta in (0.0, 0.5):
        for t in p.timesteps.tolist():
            sa, sb, sp, dc, sg = p.plan_step(t, eta).ddim
            mine = sp * ((x - sb * e) / sa) + dc * e + sg * z
            want = o.step(e, t, x, eta=eta, variance_noise=z)[0]
            torch.testing.assert_close(mine, want, rtol=1e-5, atol=1e-6)


def test_ddpm_plan_matches_oracle():
    from gm_diffusion_b200.schedulers import DDPMScheduler
    from oracle.schedulers_oracle import DDPMOracle
    p, o = DDPMScheduler(), DDPMOracle()
    p.set_timesteps(10); o.set_timesteps(10)
    assert torch.equal(p.timesteps, o.timesteps)
    g = torch.Generator().manual_seed(0)
    x, e, z = (torch.randn(8, generator=g) for _ in range(3))
    for t in p.timesteps.tolist():
        plan = p.plan_step(t)
        sa, sb, c0, c1, sg = plan.ddim
        mine = c0 * ((x - sb * e) / sa) + c1 * x + sg * z
        want = o.step(e, t, x, variance_noise=z)[0]
        torch.testing.assert_close(mine, want, rtol=1e-5, atol=1e-6)
        assert plan.needs_noise == (t > 0)


@pytest.mark.parametrize("steps", [5, 20, 50])
def test_dpmsolver_plan_matches_oracle(steps):
    """DPM-Solver++(2M) host plan (formal_improved.py:195) against the tensor-op oracle, plus the algorithm's own invariants."""
    from gm_diffusion_b200.schedulers import DPMSolverMultistepScheduler
    from oracle.schedulers_oracle import DPMSolverOracle
    p, o = DPMSolverMultistepScheduler(), DPMSolverOracle()
    p.set_timesteps(steps); o.set_timesteps(steps)
    assert torch.equal(p.timesteps, o.timesteps) and torch.equal(p.sigmas, o.sigmas)
    if steps == 50:
        assert p.timesteps[0].item() == 951 and p.timesteps[-1].item() == 20  # leading spacing over n+1 points, offset 1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(16, generator=g)
    xo = x.clone()
    prev_m = None
    for i, t in enumerate(p.timesteps.tolist()):
        e = torch.randn(16, generator=g)
        plan = p.plan_step(t)
        a_s, s_s = plan.ddim[0], plan.ddim[1]
        m0 = (x - s_s * e) / a_s
        nxt = plan.c_sample * x - plan.c_num * m0
        assert plan.plms_kind == (0 if i == 0 or i == steps - 1 else 1)
        if plan.plms_kind == 1:
            nxt = nxt - (0.5 * plan.c_num) * (plan.c_denom * (m0 - prev_m))
        prev_m, x = m0, nxt
        xo = o.step(e, t, xo)[0]
        torch.testing.assert_close(x, xo, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(x, prev_m, rtol=0, atol=0)  # final sigma 0: the last step returns the x0 prediction itself


def test_check_inputs_mirrors_reference_errors():
    from gm_diffusion_b200.pipelines._common import PipelineBase, retrieve_timesteps
    from gm_diffusion_b200.schedulers import PNDMScheduler
    pb = object.__new__(PipelineBase)
    e = torch.zeros(1, 77, 768)
    with pytest.raises(ValueError, match="divisible by 8"):
        pb.check_inputs(None, 250, 256, None, prompt_embeds=e)
    with pytest.raises(ValueError, match="Cannot forward both"):
        pb.check_inputs("a", 256, 256, None, prompt_embeds=e)
    with pytest.raises(ValueError, match="Provide either"):
        pb.check_inputs(None, 256, 256, None)
    with pytest.raises(ValueError, match="has to be of type"):
        pb.check_inputs(3, 256, 256, None)
    with pytest.raises(ValueError, match="same shape"):
        pb.check_inputs(None, 256, 256, None, prompt_embeds=e, negative_prompt_embeds=e[:, :5])
    with pytest.raises(ValueError, match="callback_steps"):
        pb.check_inputs(None, 256, 256, 0, prompt_embeds=e)
    with pytest.raises(ValueError, match="callback_on_step_end_tensor_inputs"):
        pb.check_inputs(None, 256, 256, None, prompt_embeds=e, callback_on_step_end_tensor_inputs=["nope"])
    with pytest.raises(ValueError, match="Only one of"):
        retrieve_timesteps(PNDMScheduler(), 4, None, timesteps=[1], sigmas=[1.0])
    with pytest.raises(ValueError, match="does not support custom"):
        retrieve_timesteps(PNDMScheduler(), None, None, timesteps=[1])
    ts, n = retrieve_timesteps(PNDMScheduler(), 4)
    assert n == 4 and ts.tolist() == [751, 501, 501, 251, 1]


def test_shard_range_partitions():
    from gm_diffusion_b200.dist import shard_range
    for total in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_gather_outputs_gloo_world2(tmp_path):
    """N>1 path on CPU: two gloo ranks shard 5 'images', run a stand-in per-image function, all-gather in order."""
    script = tmp_path / "w.py"
    script.write_text(f'''
import sys, torch
sys.path.insert(0, {str(ROOT)!r})
from gm_diffusion_b200 import dist as D
rank, local, world = D.init_from_env("gloo")
total = 5
a, b = D.shard_range(total, rank, world)
x = torch.arange(total, dtype=torch.float32)[a:b, None].repeat(1, 3) * 2 + 1
full = D.gather_outputs(x, total)
assert full.shape == (total, 3), full.shape
assert torch.equal(full[:, 0], torch.arange(total, dtype=torch.float32) * 2 + 1), full
assert D.max_over_ranks(float(rank), "cpu") == world - 1
print("ok", rank)
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ---- scheduler configuration: nothing the fused step kernel does not implement may be dropped silently --------------------
def test_scheduler_config_is_enforced_not_dropped():
    from types import SimpleNamespace
    from gm_diffusion_b200 import schedulers as S
    from gm_diffusion_b200.pipelines._common import as_b200_scheduler
    sd15 = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", steps_offset=1,
                set_alpha_to_one=False, skip_prk_steps=True, clip_sample=False, prediction_type="epsilon", timestep_spacing="leading")

    class DDIMScheduler:      # stands for a diffusers scheduler object: class name + .config
        def __init__(self, **kw):
            self.config = dict(sd15, **kw)

    class DDPMScheduler(DDIMScheduler):
        pass

    assert isinstance(as_b200_scheduler(DDIMScheduler()), S.DDIMScheduler)          # the SD1.5 scheduler_config.json: accepted
    assert isinstance(as_b200_scheduler(DDPMScheduler(variance_type="fixed_small")), S.DDPMScheduler)
    # diffusers LIBRARY defaults of DDIMScheduler() / DDPMScheduler(): clip_sample=True -> must refuse, not silently differ
    for cls in (DDIMScheduler, DDPMScheduler):
        with pytest.raises(NotImplementedError, match="clip_sample"):
            as_b200_scheduler(cls(clip_sample=True))
    with pytest.raises(NotImplementedError, match="timestep_spacing"):
        S.PNDMScheduler(timestep_spacing="trailing")
    with pytest.raises(NotImplementedError, match="variance_type"):
        S.DDPMScheduler(variance_type="learned_range")
    with pytest.raises(NotImplementedError, match="thresholding"):
        S.DDIMScheduler(thresholding=True)
    with pytest.raises(NotImplementedError, match="trained_betas"):
        S.DDPMScheduler(trained_betas=[0.1, 0.2])
    with pytest.raises(NotImplementedError, match="rescale_betas_zero_snr"):
        S.DDIMScheduler(rescale_betas_zero_snr=True)
    S.PNDMScheduler(clip_sample=True)     # PNDM has no clip_sample behaviour: the key is inert there, as in diffusers
    # namespace-style configs still convert
    assert isinstance(S.DDIMScheduler.from_config(SimpleNamespace(**sd15)), S.DDIMScheduler)


def test_scheduler_swap_after_construction_like_the_reference_scripts():
    """`pipeline.scheduler = DPMSolverMultistepScheduler.from_config(pipeline.scheduler.config)` (formal_improved.py:195,
    rebuttal_r2q2.py:195, rebuttal_visual.py:270) must keep working: config is a mapping AND has attributes, and every assignment
    to `.scheduler` is converted."""
    import copy
    from gm_diffusion_b200 import schedulers as S
    from gm_diffusion_b200.pipelines._common import PipelineBase
    pipe = PipelineBase.__new__(PipelineBase)      # (the constructor needs a CUDA device; the property does not)
    pipe.scheduler = S.PNDMScheduler()
    cfg = pipe.scheduler.config
    assert cfg["steps_offset"] == cfg.steps_offset == cfg.get("steps_offset") == 1 and dict(cfg)["beta_schedule"] == "scaled_linear"
    assert copy.deepcopy(cfg) == cfg
    pipe.scheduler = S.DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)      # our class
    assert isinstance(pipe.scheduler, S.DPMSolverMultistepScheduler)

    class DPMSolverMultistepScheduler:               # a diffusers-style object assigned after construction
        def __init__(self, config):
            self.config = dict(config, algorithm_type="dpmsolver++", solver_order=2, solver_type="midpoint", final_sigmas_type="zero")

        @classmethod
        def from_config(cls, config):
            assert hasattr(config, "items")           # what diffusers' ConfigMixin.from_config needs from a FrozenDict
            return cls(dict(config.items()))

    pipe.scheduler = DPMSolverMultistepScheduler.from_config(pipe.scheduler.config)
    assert isinstance(pipe.scheduler, S.DPMSolverMultistepScheduler) and hasattr(pipe.scheduler, "plan_step")
    pipe.scheduler.set_timesteps(10)
    assert len(pipe.scheduler.timesteps) == 10


def test_stale_library_is_refused(lib, monkeypatch):
    """A libgmd_b200.so built from other sources than the tree's (other struct layouts, possibly) must not load silently."""
    from gm_diffusion_b200 import build
    lib._check_fresh()                                   # the fixture just built it
    monkeypatch.setattr(build, "_digest", lambda: "0" * 64)
    with pytest.raises(RuntimeError, match="stale"):
        lib._check_fresh()
    monkeypatch.setenv("GMD_SKIP_DIGEST_CHECK", "1")
    lib._check_fresh()


def test_text_encoder_outputs_are_cached_per_prompt():
    """the default negative prompt "" (and any repeated prompt) is encoded once, not once per call."""
    from types import SimpleNamespace
    from gm_diffusion_b200.pipelines._common import PipelineBase

    class Tok:
        model_max_length = 77

        def __call__(self, texts, **kw):
            return SimpleNamespace(input_ids=torch.tensor([[len(t)] * 77 for t in texts]))

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(64, 8)
            self.calls = 0

        def forward(self, ids, **kw):
            self.calls += 1
            return (self.emb(ids),)

    pipe = PipelineBase.__new__(PipelineBase)
    pipe.tokenizer, pipe.text_encoder = Tok(), Enc()
    with torch.no_grad():
        pe1, ne1 = pipe.encode_prompt(["a cat", "a dog"], "cpu", 1, True)
        assert pipe.text_encoder.calls == 2                       # prompts, then ["", ""]
        pe2, ne2 = pipe.encode_prompt(["a dog", "a cat"], "cpu", 1, True)
    assert pipe.text_encoder.calls == 2                           # everything came from the cache
    assert torch.equal(pe2[0], pe1[1]) and torch.equal(pe2[1], pe1[0]) and torch.equal(ne1, ne2) and ne1.shape == (2, 77, 8)


def test_scratch_slots_keep_concurrent_streams_apart():
    """The library's static scratch (split-K workspace, GroupNorm accumulator arena) is keyed by (device, slot): the loop graph runs
    the GM branch on a side stream under `ops.scratch_slot(1)` beside the SDR UNet on slot 0 (stable_diffusion_dual_unet.py
    denoise_loop_two_streams).  Host logic only: keys nest, restore on exit (also on an exception) and never alias across slots."""
    from gm_diffusion_b200 import ops
    assert ops._scratch_key("cuda:0") == (0, 0) and ops._scratch_key("cuda:3") == (3, 0)
    with ops.scratch_slot(1):
        assert ops._scratch_key("cuda:0") == (0, 1)
        with ops.scratch_slot(2):
            assert ops._scratch_key("cuda:0") == (0, 2)
        assert ops._scratch_key("cuda:0") == (0, 1)
        assert ops.gn_arena("cuda:0").key == (0, 1)
    assert ops._scratch_key("cuda:0") == (0, 0)
    with pytest.raises(RuntimeError):
        with ops.scratch_slot(5):
            raise RuntimeError("boom")
    assert ops._scratch_key("cuda:0") == (0, 0)
