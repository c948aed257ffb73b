"""Token-stream GEMMs of the transformer blocks, timed the way they run inside the step: from a CUDA graph, cycling through buffer sets
larger than L2 (so activations come from HBM like they do behind a 1.7 GB weight stream), tiled weights.
  python profiles/bench_token_gemm.py            (all shapes)      python profiles/bench_token_gemm.py plain 1   (one shape once, for ncu)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import os
from gm_diffusion_b200 import _lib, ops
if os.environ.get('GMD_AB_LIB'):   # A/B runs against an alternative build of the library
    _lib.LIB_PATH = Path(os.environ['GMD_AB_LIB']).resolve()

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16
SHAPES = {  # name: (M, N, K, kind)
    "plain_320": (65536, 320, 320, "plain"), "qkv_320": (65536, 960, 320, "plain"), "res32_320": (65536, 320, 320, "res32"),
    "geglu_320": (65536, 2560, 320, "geglu"), "ff2_320": (65536, 320, 1280, "res32"),
    "plain_640": (16384, 640, 640, "plain"), "res32_640": (16384, 640, 640, "res32"), "geglu_640": (16384, 5120, 640, "geglu"),
    "plain_1280": (4096, 1280, 1280, "plain"), "res32_1280": (4096, 1280, 1280, "res32"),
    # LayerNorm folding: producer (fp32 stream + bf16 copy + row statistics), consumer (normalise in the epilogue)
    "res32_320_lnout": (65536, 320, 320, "res32+lnout"), "qkv_320_lnin": (65536, 960, 320, "plain+lnin"), "plain_320_lnin": (65536, 320, 320, "plain+lnin"),
    "res32_640_lnout": (16384, 640, 640, "res32+lnout"), "qkv_640_lnin": (16384, 1920, 640, "plain+lnin"),
}
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for name, (M, N, K, kind) in SHAPES.items():
    if only and only != name:
        continue
    lnout, lnin = kind.endswith("+lnout"), kind.endswith("+lnin")
    kind = kind.split("+")[0]
    n_out = N // 2 if kind == "geglu" else N
    per_set = M * K * 2 + M * n_out * (4 if kind == "res32" else 2) + (M * N * 4 if kind == "res32" else 0)
    nsets = max(2, int(400e6 // per_set) + 1)
    a = [torch.randn(M, K, device=dev, generator=g).to(bf) for _ in range(nsets)]
    w = ops.tile_weight((torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(bf), geglu=(kind == "geglu"))
    b = torch.randn(N, device=dev, generator=g)
    res = [torch.randn(M, N, device=dev, generator=g) for _ in range(nsets)] if kind == "res32" else None
    out = [torch.empty(M, n_out, device=dev, dtype=torch.float32 if kind == "res32" else bf) for _ in range(nsets)]

    ln_s = torch.zeros(M, 2, dtype=torch.int64, device=dev) + (1 << 24)
    ln_c = torch.randn(N, device=dev, generator=g)

    def call(i):
        if kind == "plain":
            ops.gemm(a[i], w, bias=b, out=out[i], ln_in=(ln_s, ln_c, 1e-5) if lnin else None)
        elif kind == "res32":
            ops.gemm(a[i], w, bias=b, residual=res[i], out=out[i], out_f32=True, ln_out=lnout)
        else:
            ops.gemm(a[i], w, bias=b, geglu=True, out=out[i])
    for i in range(nsets):
        call(i)
    torch.cuda.synchronize()
    if reps == 1:
        call(0); torch.cuda.synchronize()
        continue
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for r in range(reps):
            call(r % nsets)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1000)
    us = sorted(ts)[2]
    byts = M * K * 2 + N * K * 2 + M * n_out * (4 if kind == "res32" else 2) + (M * N * 4 if kind == "res32" else 0)
    print(f"{name:12s} M={M} N={N} K={K} {kind:6s}: {us:7.1f} us  {2.0 * M * N * K / us * 1e-6:7.1f} TFLOP/s  {byts / us * 1e-6:6.2f} TB/s of algorithmic bytes ({byts / 1e6:.0f} MB, {nsets} buffer sets)")
