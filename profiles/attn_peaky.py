"""Accuracy of the self-attention kernels on peaky softmax rows (large logits, maxima that grow along the key axis): rel-L2 against fp32 torch.
The table in attn_ab_r02c.txt; `GMD_AB_LIB` selects an A/B build."""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import _lib, ops
if os.environ.get('GMD_AB_LIB'):
    _lib.LIB_PATH = Path(os.environ['GMD_AB_LIB']).resolve()
bf = torch.bfloat16
for d, N, ramp, seed, amp in [(40, 512, False, 2, 6), (40, 512, False, 554, 6), (40, 512, False, 7, 6), (80, 512, False, 594, 6), (40, 1024, True, 1066, 6), (80, 768, True, 850, 6), (40, 512, False, 554, 3), (40, 512, False, 554, 1.5), (40,4096,False,1,1)]:
    g = torch.Generator().manual_seed(seed)
    B, H = 1, 8
    q = (torch.randn(B, N, H * d, generator=g) * amp).to(bf).cuda()
    k = torch.randn(B, N, H * d, generator=g) * amp
    if ramp:
        k = k * torch.linspace(0.05, 2.5, N).view(1, N, 1)
    k = k.to(bf).cuda()
    v = torch.randn(B, N, H * d, generator=g).to(bf).cuda()
    got = ops.attention(q, k, v, H).float()
    qh, kh, vh = (t.float().reshape(B, -1, H, d).transpose(1, 2) for t in (q, k, v))
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, N, H * d)
    err = (got - ref).pow(2).sum(-1).sqrt() / ref.pow(2).sum(-1).sqrt().clamp_min(1e-6)   # per row
    bad = (err[0] > 0.05).nonzero().flatten()
    print(f"d={d} N={N} ramp={ramp} seed={seed} amp={amp}: rel-L2 {((got-ref).pow(2).sum()/ref.pow(2).sum()).sqrt().item():.3e} finite={bool(torch.isfinite(got).all())} bad rows {bad.numel()} first {bad[:8].tolist()}", flush=True)
