"""CPU oracle for the GM-Diffusion Stage-3 hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this package; the product (`gm_diffusion_b200`) never does and has no CPU path of its own.

Pinning status:
  * `tone_mapping_oracle`  — PINNED: checked against the reference's own
    `gm_diffusion/stage1/tone_mapping.py` executed by file path in the build container; the resulting
    input/output vectors are committed under `tests/golden/tm_*.npz` (generator: `oracle/make_golden.py`).
  * `schedulers_oracle`, `unet_oracle`, `vae_oracle`, `pipeline_oracle` — PARITY UNPINNED by the
    reference: the arithmetic lives in third-party `diffusers` (>=0.33, unpinned, not vendored, not
    installable offline) and the reference has no tests or golden vectors.  These modules restate the
    published diffusers algorithms and the reference's own loop
    (`stable_diffusion_dual_unet.py:1040-1093`, `stable_diffusion_gm.py:1040-1071`); they are pinned by
    analytic invariants instead (parameter count 859 520 964 / 859 532 484, PNDM timestep literals,
    PLMS == DDIM identity, constant-eps closed form).
"""
