"""Per-call device time of one VAE decode (8 latents 64x64 -> 8 images 512x512), grouped by op shape (eager CUDA events: kernels shorter
than ~10 us are overstated by the launch gap).  Run on a B200:  python profiles/layer_times_vae.py > gpurun_out/layer_times_vae.txt"""
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
names = ["gemm", "conv2d", "attention", "groupnorm_silu", "layernorm", "softmax_rows"]
orig = {n: getattr(ops, n) for n in names}
rec = []

def desc(n, a, k, out):
    if n == "gemm":
        x, w = a[0], a[1]
        fl = 2.0 * x.shape[-2] * w.shape[-2] * x.shape[-1] * (x.shape[0] if x.dim() == 3 else 1)
        return f"gemm {tuple(x.shape)} x {tuple(w.shape)}" + (" f32out" if out.dtype == torch.float32 else ""), fl
    if n == "conv2d":
        x, w = a[0], a[1]
        fl = 2.0 * (out.numel() // out.shape[-1]) * a[2] * w.shape[1]
        return f"conv{k.get('ksize', 3)} {tuple(x.shape[:3])} K={w.shape[1]} Cout={a[2]}" + (" up" if k.get("upsample") else "") + (" f32out" if out.dtype == torch.float32 else ""), fl
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
x = torch.randn(B * 64 * 64, 4, device=dev, generator=g)
for it in range(2):
    for n in names: setattr(ops, n, wrap(n))
    rec.clear(); torch.cuda.synchronize()
    pipe.vae.decode_px(x, B, 64, 64)
    torch.cuda.synchronize()
    for n in names: setattr(ops, n, orig[n])
agg = defaultdict(lambda: [0.0, 0.0, 0])
for d, fl, e0, e1 in rec:
    a = agg[d]; a[0] += e0.elapsed_time(e1); a[1] += fl; a[2] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot:.2f} ms over {len(rec)} calls")
for d, (ms, fl, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{n:3d} {fl / ms * 1e-9 if ms else 0:8.1f} TF/s  {d}")
