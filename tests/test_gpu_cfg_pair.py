"""CFG-pair mode (optional): two ranks split the classifier-free-guidance halves of the SDR UNet and the images of
the GM UNet, exchange eps once per step, and must reproduce the single-GPU latents bit for bit.  With two GPUs the ranks sit on
cuda:0 / cuda:1 and exchange over NCCL; on a one-GPU box both ranks share cuda:0 and exchange through gloo (host-staged: kernels of
the two processes never wait on one another) — the pair logic, the split and the bit-identity claim are the same.
Also runnable directly: `python tests/test_gpu_cfg_pair.py`."""
import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _worker(rank: int, port: int, B: int):
    import torch.distributed as dist
    import gm_diffusion_b200 as G
    from gm_diffusion_b200 import random_init as R
    two_gpus = torch.cuda.device_count() >= 2
    idx = rank if two_gpus else 0
    torch.cuda.set_device(idx)
    dev = torch.device("cuda", idx)
    dist.init_process_group("nccl" if two_gpus else "gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=2)
    try:
        sd4 = R.sd15_unet_state_dict(4, seed=0, device=dev)
        pipe = G.StableDiffusionDualUNetPipeline(vae=None, text_encoder=None, tokenizer=None, unet=G.B200UNet(sd4, device=dev),
                                                 gm_unet=G.B200UNet(R.widen_conv_in_state_dict(sd4), device=dev), scheduler=G.PNDMScheduler(), device=dev)
        del sd4
        g = torch.Generator().manual_seed(1)
        pe, ne = torch.randn(B, 77, 768, generator=g).to(dev), torch.randn(B, 77, 768, generator=g).to(dev)
        lat = torch.randn(B, 4, 32, 32, generator=g).to(dev)
        kw = dict(prompt_embeds=pe, negative_prompt_embeds=ne, height=256, width=256, num_inference_steps=3, guidance_scale=7.5, output_type="latent")
        for graphs in (False, True):
            pipe.use_cuda_graph = graphs
            pipe.cfg_pair = None
            s0, g0 = pipe(latents=lat.clone(), **kw)                     # every rank alone: the single-GPU result
            pipe.enable_cfg_pair()
            s1, g1 = pipe(latents=lat.clone(), **kw)
            assert torch.equal(s0, s1) and torch.equal(g0, g1), f"rank {rank} graphs={graphs}: CFG-pair result differs from the single-GPU result"
        # and the two ranks hold the same tensors
        mine = s1.contiguous() if two_gpus else s1.cpu().contiguous()
        both = [torch.empty_like(mine) for _ in range(2)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [2, 3])
def test_cfg_pair_matches_single_gpu(B):
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(29650 + B, B), nprocs=2, join=True)


if __name__ == "__main__":
    import torch.multiprocessing as mp
    for B in (2, 3):
        mp.spawn(_worker, args=(29650 + B, B), nprocs=2, join=True)
        print(f"cfg-pair B={B}: identical to the single-GPU result on both ranks")
