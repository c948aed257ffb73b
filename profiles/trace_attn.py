"""Debug timeline of attn2_kernel (library built with -DGMD_ATTN2_TRACE): clock64 stamps of CTA (0,0,0) per key tile.
slots: 0 K_j TMA issued | 1 MMA thread saw k_full(j) | 2 S_j issued | 3 MMA thread saw p_full(j) | 4 saw v_full(j) | 5 P V_j issued |
6 softmax saw s_full(j) | 7 softmax arrived on p_full(j)"""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from gm_diffusion_b200 import _lib, ops
_lib.LIB_PATH = Path(os.environ["GMD_AB_LIB"]).resolve()
B, N, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 8)
torch.cuda.synchronize()
T = N // 64
buf = (ctypes.c_longlong * (8 * 1024))()
rc = _lib.lib().gmd_attn_trace_dump(buf, 8 * 1024)
assert rc == 0, rc
dc, dt = buf[1001 * 8] - buf[1000 * 8], buf[1001 * 8 + 1] - buf[1000 * 8 + 1]
if dt > 0:
    print(f"CTA (0,0,0): {dc} cycles in {dt} ns between its first and last tile -> SM clock {dc / dt:.3f} GHz")
t0 = min(buf[i] for i in range(8 * T) if buf[i] > 0)
names = ["K_issue", "k_full", "S_issue", "p_full", "v_full", "PV_issue", "s_full", "p_arrive"]
print("tile " + " ".join(f"{n:>9}" for n in names))
for j in range(T):
    print(f"{j:4d} " + " ".join(f"{buf[j * 8 + s] - t0:9d}" for s in range(8)))
