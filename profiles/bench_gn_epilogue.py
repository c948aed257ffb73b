"""What do the GroupNorm statistics cost in the producer's epilogue, and what does the consumer save?  Per-call time inside a CUDA
graph (24 calls on rotating buffers, replayed), with and without gn_stats, at the denoise step's level-0 / level-1 shapes."""
import math, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import ops
bf = torch.bfloat16
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)

def graph_us(fn, n=24, reps=5):
    with ops.gn_arena(dev, None):
        fn(0); fn(1)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        with ops.gn_arena(dev, None):
            for i in range(n):
                fn(i)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (reps * n)

ws = ops.gn_workspace(64, 32, dev)
for (N, H, C) in ((16, 64, 320), (16, 32, 640), (16, 16, 1280), (8, 64, 320)):
    NB = 4
    xs = [torch.randn(N, H, H, C, device=dev, generator=g).to(bf) for _ in range(NB)]
    res = [torch.randn(N, H, H, C, device=dev, generator=g).to(bf) for _ in range(NB)]
    w3 = ops.pack_conv_weight_tiled((torch.randn(C, C, 3, 3, device=dev, generator=g) / math.sqrt(9 * C)))
    w1 = ops.tile_weight((torch.randn(C, C, device=dev, generator=g) / math.sqrt(C)))
    b = torch.randn(C, device=dev, generator=g)
    gam, bet = torch.randn(C, device=dev, generator=g), torch.randn(C, device=dev, generator=g)
    M = N * H * H
    row = {}
    row["conv3+res"] = graph_us(lambda i: ops.conv2d(xs[i % NB], w3, C, bias=b, residual=res[i % NB]))
    row["conv3+res+stats"] = graph_us(lambda i: ops.conv2d(xs[i % NB], w3, C, bias=b, residual=res[i % NB], gn_stats=True))
    row["conv3 f32out"] = graph_us(lambda i: ops.conv2d(xs[i % NB], w3, C, bias=b, out_f32=True))
    row["conv3 f32out+stats"] = graph_us(lambda i: ops.conv2d(xs[i % NB], w3, C, bias=b, out_f32=True, gn_stats=True))
    row["gemm K=C +res"] = graph_us(lambda i: ops.gemm(xs[i % NB].view(M, C), w1, bias=b, residual=res[i % NB].view(M, C)))
    row["gemm K=C +res+stats"] = graph_us(lambda i: ops.gemm(xs[i % NB].view(M, C), w1, bias=b, residual=res[i % NB].view(M, C), gn_rows_per_sample=H * H))
    _, sums = ops.conv2d(xs[0], w3, C, bias=b, gn_stats=True)
    y32 = [x.float() for x in xs]
    row["GN one-pass bf16"] = graph_us(lambda i: ops.groupnorm_silu(xs[i % NB], gam, bet, stats_ws=ws))
    row["GN apply (epilogue stats) bf16"] = graph_us(lambda i: ops.groupnorm_silu(xs[i % NB], gam, bet, sums=sums))
    row["GN one-pass fp32 in"] = graph_us(lambda i: ops.groupnorm_silu(y32[i % NB], gam, bet, stats_ws=ws))
    row["GN apply (epilogue stats) fp32 in"] = graph_us(lambda i: ops.groupnorm_silu(y32[i % NB], gam, bet, sums=sums))
    print(f"({N}, {H}, {H}, {C}):  " + "  ".join(f"{k} {v:.1f}" for k, v in row.items()), flush=True)
