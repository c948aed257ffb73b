"""GPU diagnostic (not a pytest): op-by-op comparison of two B200UNet runs (batch 3 vs its middle sample alone) and
run-to-run determinism.  Usage: python tests/diag_unet.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import B200UNet, ops
from oracle.unet_oracle import UNet2DConditionOracle

torch.manual_seed(0)
u4 = UNet2DConditionOracle(4).eval().cuda()
m = B200UNet.from_module(u4)
g = torch.Generator().manual_seed(7)
x = torch.randn(3, 4, 32, 32, generator=g).cuda()
ctx = torch.randn(3, 77, 768, generator=g).cuda()

rec = []
names = ["gemm", "conv2d", "groupnorm_silu", "layernorm", "attention"]
orig = {n: getattr(ops, n) for n in names}
def wrap(n):
    def f(*a, **k):
        o = orig[n](*a, **k)
        rec.append((n, o.detach().clone()))
        return o
    return f
for n in names:
    setattr(ops, n, wrap(n))

def run(xx, cc):
    rec.clear()
    out = m.forward_nchw(xx, 500, cc)
    torch.cuda.synchronize()
    return out, list(rec)

o1, r1 = run(x, ctx)
o2, r2 = run(x, ctx)
print("run-to-run max abs diff", float((o1 - o2).abs().max()))
for i, ((n, a), (_, b)) in enumerate(zip(r1, r2)):
    d = float((a.float() - b.float()).abs().max())
    if d > 0:
        print("  nondeterministic op", i, n, tuple(a.shape), d); break
o3, r3 = run(x[1:2], ctx[1:2])
print("batch3[1] vs single rel", float((o1[1:2] - o3).norm() / o3.norm()))
B = 3
for i, ((n, a), (_, b)) in enumerate(zip(r1, r3)):
    if a.shape[0] == B * b.shape[0]:
        k = b.shape[0]
        sl = a[k:2 * k]
    elif a.dim() == 3 and a.shape[0] == B:
        sl = a[1:2]
    else:
        print("  skip", i, n, tuple(a.shape), tuple(b.shape)); continue
    r = float((sl.float() - b.float()).norm() / (b.float().norm() + 1e-20))
    if r > 1e-4 or i < 6:
        print(f"  op {i:3d} {n:15s} {tuple(a.shape)} rel {r:.3e}")
    if r > 1e-3:
        break

# ---- noise floor: the oracle itself executed in bf16 by torch (what the reference would run) vs fp32 ----
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import copy
with torch.no_grad():
    for hw, t in [(32, 981), (64, 501), (16, 741)]:
        g = torch.Generator().manual_seed(100 + hw + t)
        xx = torch.randn(2, 4, hw, hw, generator=g).cuda(); cc = torch.randn(2, 77, 768, generator=g).cuda()
        want = u4(xx, t, encoder_hidden_states=cc)
        ub = copy.deepcopy(u4).bfloat16()
        tb = ub(xx.bfloat16(), t, encoder_hidden_states=cc.bfloat16()).float()
        for n in names: setattr(ops, n, orig[n])
        mine = m.forward_nchw(xx, t, cc)
        rl = lambda a, b: float((a - b).norm() / b.norm())
        print(f"hw={hw} t={t}: torch-bf16 vs fp32 {rl(tb, want):.3e} | b200 vs fp32 {rl(mine, want):.3e} | b200 vs torch-bf16 {rl(mine, tb):.3e} | eps rms {float(want.pow(2).mean().sqrt()):.3f}")
        del ub
