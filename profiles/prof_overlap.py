"""Would running the SDR UNet of step t+1 concurrently with the GM UNet of step t (independent chains) pay?  Times the two CUDA graphs
back to back on one stream vs concurrently on two streams."""
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
unet_in = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
gm_in = torch.randn(B, h, w, 8, device=dev, generator=g).to(torch.bfloat16)
eps_s = torch.empty(2 * B, h, w, 4, device=dev); eps_g = torch.empty(B, h, w, 4, device=dev)

def fs(): pipe.unet.forward(unet_in, tb_s, kv_s, out=eps_s, cfg_shared=True)
def fg(): pipe.gm_unet.forward(gm_in, tb_g, kv_g, out=eps_g)
for f in (fs, fg):
    f(); f()
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
graphs = []
for f, st in ((fs, s1), (fg, s2)):
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        f()
    graphs.append(gr)
torch.cuda.synchronize()

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def seq():
    graphs[0].replay(); graphs[1].replay()
def conc():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): graphs[0].replay()
    with torch.cuda.stream(s2): graphs[1].replay()
    cur.wait_stream(s1); cur.wait_stream(s2)
print(f"SDR graph alone {timed(lambda: graphs[0].replay()):.2f} ms, GM graph alone {timed(lambda: graphs[1].replay()):.2f} ms")
print(f"sequential {timed(seq):.2f} ms   concurrent on two streams {timed(conc):.2f} ms")
