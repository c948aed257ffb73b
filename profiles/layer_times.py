"""Per-call device time of one eager denoise step (SDR UNet on 2B + GM UNet on B, 512x512, B=8), grouped by op shape.
Run on a B200:  python profiles/layer_times.py > gpurun_out/layer_times.txt"""
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from gm_diffusion_b200 import ops

dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
B = 8
names = ["gemm", "conv2d", "attention", "groupnorm_silu", "layernorm"]
orig = {n: getattr(ops, n) for n in names}
rec = []

def desc(n, a, k, out):
    if n == "gemm":
        x, w = a[0], a[1]
        fl = 2.0 * x.shape[-2] * w.shape[-2] * x.shape[-1]
        return f"gemm M={x.shape[-2]} N={w.shape[-2]} K={x.shape[-1]}" + (" geglu" if k.get("geglu") else "") + (" f32out" if out.dtype == torch.float32 else ""), fl
    if n == "conv2d":
        x, w = a[0], a[1]
        fl = 2.0 * (out.numel() // out.shape[-1]) * a[2] * w.shape[1]
        tag = f"conv{k.get('ksize', 3)} {tuple(x.shape[:3])} K={w.shape[1]} Cout={a[2]}" + (" s2" if k.get("stride", 1) == 2 else "") + (" up" if k.get("upsample") else "") + (" 2src" if k.get("x1") is not None else "")
        return tag, fl
    if n == "attention":
        q, kk = a[0], a[1]
        return f"attn B={q.shape[0]} Nq={q.shape[1]} Nk={kk.shape[1]} C={q.shape[2]}", 4.0 * q.shape[0] * q.shape[1] * kk.shape[1] * q.shape[2]
    x = a[0]
    return f"{n} {tuple(x.shape)} {str(x.dtype)[6:]}", 0.0

def wrap(n):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = orig[n](*a, **k); e1.record()
        d, fl = desc(n, a, k, out[0] if isinstance(out, tuple) else out)
        rec.append((d, fl, e0, e1))
        return out
    return f

g = torch.Generator(device=dev).manual_seed(5)
ctx2 = torch.randn(2 * B, 77, 768, device=dev, generator=g)
kv_s, kv_g = pipe.unet.project_context(ctx2), pipe.gm_unet.project_context(ctx2[:B])
tb_s, tb_g = pipe.unet.timestep_table([501]), pipe.gm_unet.timestep_table([501])
xs = torch.randn(B, 64, 64, 8, device=dev, generator=g).to(torch.bfloat16)
xg = torch.randn(B, 64, 64, 8, device=dev, generator=g).to(torch.bfloat16)
for it in range(2):
    for n in names: setattr(ops, n, wrap(n))
    rec.clear(); torch.cuda.synchronize()
    pipe.unet.forward(xs, tb_s, kv_s, cfg_shared=True); pipe.gm_unet.forward(xg, tb_g, kv_g)
    torch.cuda.synchronize()
    for n in names: setattr(ops, n, orig[n])
agg = defaultdict(lambda: [0.0, 0.0, 0])
for d, fl, e0, e1 in rec:
    a = agg[d]; a[0] += e0.elapsed_time(e1); a[1] += fl; a[2] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot:.2f} ms over {len(rec)} calls")
for d, (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n:3d}  {tf:7.1f} TF/s  {d}")
