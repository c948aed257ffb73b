"""Latency of one 512x512 / 50-step dual-branch generation on ONE GPU vs a CFG pair (two GPUs working on the same images).
torchrun --nproc-per-node 2 profiles/bench_cfg_pair.py"""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import bench
from gm_diffusion_b200 import dist as D

rank, local, world = D.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
pipe = bench.build_pipeline(dev)
for B in (1, 2, 8):
    g = torch.Generator().manual_seed(B)
    pe, ne = torch.randn(B, 77, 768, generator=g).to(dev), torch.randn(B, 77, 768, generator=g).to(dev)
    lat = torch.randn(B, 4, 64, 64, generator=g).to(dev)
    kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, output_type="latent")
    res = {}
    for mode in ("single", "pair"):
        pipe.cfg_pair = None
        if mode == "pair":
            pipe.enable_cfg_pair()
        for _ in range(2):
            pipe(latents=lat.clone(), **kw)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pipe(latents=lat.clone(), **kw)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1), out)
    same = torch.equal(res["single"][1][0], res["pair"][1][0]) and torch.equal(res["single"][1][1], res["pair"][1][1])
    if rank == 0:
        print(f"B={B}: one GPU {res['single'][0]:8.1f} ms   CFG pair (2 GPUs) {res['pair'][0]:8.1f} ms   speed-up {res['single'][0] / res['pair'][0]:.2f}x   identical: {same}", file=sys.stderr)
dist.barrier()
dist.destroy_process_group()
