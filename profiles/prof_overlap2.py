"""Upper bound of two-stream overlap: the SDR UNet forward (16 samples) and the GM UNet forward (8 samples) of a denoise step as two
CUDA graphs, replayed back to back on one stream vs concurrently on two streams (timing only: the two forwards share scratch here)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
dev = torch.device("cuda:0")
pipe = bench.build_pipeline(dev)
B, h, w = 8, 64, 64
g = torch.Generator(device=dev).manual_seed(5)
ctx2 = torch.randn(2 * B, 77, 768, device=dev, generator=g)
kv_s, kv_g = pipe.unet.project_context(ctx2), pipe.gm_unet.project_context(ctx2[B:])
tb_s, tb_g = pipe.unet.timestep_table([981]), pipe.gm_unet.timestep_table([981])
xs = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
xg = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
eps_s = torch.empty(2 * B, h, w, 4, device=dev); eps_g = torch.empty(B, h, w, 4, device=dev)
fa = lambda: pipe.unet.forward(xs, tb_s, kv_s, out=eps_s, cfg_shared=True)
fb = lambda: pipe.gm_unet.forward(xg, tb_g, kv_g, out=eps_g)
fa(); fb(); torch.cuda.synchronize()
ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
with torch.cuda.graph(ga):
    fa()
with torch.cuda.graph(gb):
    fb()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def seq():
    ga.replay(); gb.replay()
def par():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1):
        ga.replay()
    with torch.cuda.stream(s2):
        gb.replay()
    cur.wait_stream(s1); cur.wait_stream(s2)
print(f"SDR forward alone {timeit(ga.replay):.3f} ms, GM forward alone {timeit(gb.replay):.3f} ms")
print(f"sequential {timeit(seq):.3f} ms   two streams {timeit(par):.3f} ms")
