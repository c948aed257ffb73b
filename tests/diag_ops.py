"""GPU diagnostic (not a pytest): run the VAE decoder / UNet with every gemm / conv2d call checked against a torch fp32
computation of the same op on the same inputs; prints the first calls whose relative error is large."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
from gm_diffusion_b200 import ops, B200VaeDecoder, random_init as R

orig_gemm, orig_conv = ops.gemm, ops.conv2d
bad = []
def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))
def gemm(a, w, **k):
    out = orig_gemm(a, w, **k)
    ref = torch.matmul(a.float(), w.float().transpose(-1, -2))
    if k.get("alpha") is not None: ref = ref * k["alpha"]
    if k.get("bias") is not None and not k.get("geglu"): ref = ref + k["bias"]
    if k.get("geglu"):
        Nn = w.shape[-2]; half = 80
        full = ref + 0
        # undo the tile interleave: [value half-tile | gate half-tile] per 160 rows
        v = full.reshape(*full.shape[:-1], Nn // 160, 2, half)
        b = k["bias"]; inner = Nn // 2
        val = v[..., 0, :].reshape(*full.shape[:-1], inner) + b[:inner]
        gate = v[..., 1, :].reshape(*full.shape[:-1], inner) + b[inner:]
        ref = val * F.gelu(gate)
    if k.get("row_bias") is not None: ref = ref + k["row_bias"].repeat_interleave(k["rows_per_sample"], 0)
    if k.get("residual") is not None: ref = ref + k["residual"].float()
    r = rel(out, ref)
    tag = f"gemm a{tuple(a.shape)} w{tuple(w.shape)} strides a{a.stride()} w{w.stride()} out{tuple(out.shape)} {out.dtype} keys={[x for x in k if k[x] is not None and x != 'out']}"
    if not (r < 2e-2): bad.append((r, tag)); print("BAD", r, tag, "finite", bool(torch.isfinite(out).all()))
    return out
def conv2d(x, w, cout, **k):
    out = orig_conv(x, w, cout, **k)
    ks = k.get("ksize", 3)
    xin = torch.cat([x, k["x1"]], -1) if k.get("x1") is not None else x
    cin = xin.shape[-1]
    w4 = w.float().reshape(w.shape[0], ks, ks, cin).permute(0, 3, 1, 2)
    xx = xin.float().permute(0, 3, 1, 2)
    if k.get("upsample"): xx = F.interpolate(xx, scale_factor=2.0, mode="nearest")
    ref = F.conv2d(xx, w4, k.get("bias"), stride=k.get("stride", 1), padding=ks // 2).permute(0, 2, 3, 1)[..., :cout]
    if k.get("row_bias") is not None: ref = ref + k["row_bias"][:, None, None, :]
    if k.get("residual") is not None: ref = ref + k["residual"].float()
    r = rel(out, ref)
    tag = f"conv x{tuple(x.shape)} w{tuple(w.shape)} cout={cout} {out.dtype} keys={[a for a in k if k[a] is not None and a != 'out']}"
    if not (r < 2e-2): bad.append((r, tag)); print("BAD", r, tag, "finite", bool(torch.isfinite(out).all()))
    return out
ops.gemm, ops.conv2d = gemm, conv2d
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
which = sys.argv[1] if len(sys.argv) > 1 else "vae"
if which == "vae":
    vae = B200VaeDecoder(R.sd_vae_decoder_state_dict(seed=4, device="cuda"))
    z = torch.randn(2, 4, 16, 16, device="cuda")
    img = vae.decode(z)
    print("vae out finite", bool(torch.isfinite(img).all()), tuple(img.shape))
else:
    from gm_diffusion_b200 import B200UNet
    u = B200UNet(R.sd15_unet_state_dict(4, 0, "cuda"))
    x = torch.randn(2, 4, 32, 32, device="cuda"); ctx = torch.randn(2, 77, 768, device="cuda")
    e = u.forward_nchw(x, 500, ctx)
    print("unet out finite", bool(torch.isfinite(e).all()))
print("bad calls:", len(bad))
