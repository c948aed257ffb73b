"""Where does a short-K token GEMM spend its time?  M = 65536, N = 320: K sweep (main loop share) x epilogue variants."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
M = 65536
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n): fn()
    graph.replay(); torch.cuda.synchronize()
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for N in (320, 960):
    for K in (64, 128, 320, 640, 1280):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        w = ops.tile_weight((torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16))
        b = torch.randn(N, device="cuda", generator=g)
        res = torch.randn(M, N, device="cuda", generator=g)
        out16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
        t = {"bias": bench(lambda: ops.gemm(a, w, bias=b, out=out16)), "nobias": bench(lambda: ops.gemm(a, w, out=out16)),
             "f32out": bench(lambda: ops.gemm(a, w, bias=b, out=out32)), "f32out+res": bench(lambda: ops.gemm(a, w, bias=b, residual=res, out=out32))}
        print(f"N={N} K={K:5d}: " + "  ".join(f"{k} {v:6.1f} us" for k, v in t.items()), f"  (MMA floor {2 * M * N * K / 1.6538e15 * 1e6:5.1f} us)")
