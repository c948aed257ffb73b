"""tcgen05 GEMM / implicit-GEMM conv / flash attention / norms (C-ABI) vs plain PyTorch fp32 references of the
same op on the same bf16-rounded inputs.  Tolerance: bf16 output rounding (rel 2^-8) + fp32 accumulation order."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf = torch.bfloat16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-20))


def check(got, want, tol=6e-3, what=""):
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (got.shape, want.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    r = rel_l2(got, want)
    mx = float((got - want).abs().max())
    assert r < tol, f"{what}: rel-L2 {r:.3e} (max abs {mx:.3e}, ref rms {float(want.pow(2).mean().sqrt()):.3e})"
    # elementwise: bf16 rounding of the output + small absolute slack
    scale = float(want.abs().max()) + 1e-6
    assert mx <= 2.5e-2 * scale, f"{what}: max abs err {mx:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("M,N,K", [(128, 160, 64), (256, 320, 320), (1000, 320, 768), (77 * 2, 640, 768), (4096, 1280, 1280),
                                   (130, 128, 72), (64, 32, 64), (300, 512, 512), (16, 2560, 1280)])
def test_gemm_plain(M, N, K):
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(bf).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(bf).cuda()
    check(ops.gemm(a, w), a.float() @ w.float().t(), what=f"gemm {M}x{N}x{K}")
    check(ops.gemm(a, ops.tile_weight(w)), a.float() @ w.float().t(), what=f"gemm tiled+swizzled-W {M}x{N}x{K}")
    assert torch.equal(ops.gemm(a, ops.tile_weight(w, swizzle=False)), ops.gemm(a, ops.tile_weight(w))), "tiled vs tiled+pre-swizzled weights"


def test_gemm_epilogues():
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(0)
    M, N, K, S = 512, 640, 320, 128
    a = torch.randn(M, K, generator=g).to(bf).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(bf).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).to(bf).cuda()
    rb = torch.randn(M // S, N, generator=g).cuda()
    ref = a.float() @ w.float().t()
    check(ops.gemm(a, w, bias=bias), ref + bias, what="bias")
    check(ops.gemm(a, w, bias=bias, residual=res), ref + bias + res.float(), what="bias+residual")
    res32 = torch.randn(M, N, generator=g).cuda()
    check(ops.gemm(a, w, bias=bias, residual=res32, out_f32=True), ref + bias + res32, tol=1e-3, what="fp32 residual stream")
    check(ops.gemm(a, w, bias=bias, row_bias=rb, rows_per_sample=S), ref + bias + rb.repeat_interleave(S, 0), what="row_bias")
    out = ops.gemm(a, w, bias=bias, out_f32=True)
    assert out.dtype == torch.float32
    check(out, ref + bias, tol=1e-3, what="fp32 out")
    check(ops.gemm(a, w, alpha=0.125), ref * 0.125, what="alpha")
    # strided views (QKV-style column slices) as A, and as output
    big = torch.randn(M, 3 * K, generator=g).to(bf).cuda()
    check(ops.gemm(big[:, K:2 * K], w), big[:, K:2 * K].float() @ w.float().t(), what="strided A")


def test_gemm_splitk_paths():
    """Few output tiles + long K -> split-K with the deterministic finalize kernel (bias / row bias / residual / dtypes)."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(11)
    M, N, K = 512, 1280, 5120
    a = torch.randn(M, K, generator=g).to(bf).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(bf).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res32 = torch.randn(M, N, generator=g).cuda()
    res16 = torch.randn(M, N, generator=g).to(bf).cuda()
    rb = torch.randn(4, N, generator=g).cuda()
    ref = a.float() @ w.float().t() + bias
    from gm_diffusion_b200 import _lib as L
    L.reset_launch_count()
    check(ops.gemm(a, w, bias=bias, splitk=True), ref, what="split-K bias")
    assert L.launch_count() == 2, "split-K = main kernel + finalize kernel"
    check(ops.gemm(a, w, bias=bias, residual=res32, out_f32=True, splitk=True), ref + res32, tol=1e-3, what="split-K fp32 residual")
    check(ops.gemm(a, w, bias=bias, residual=res16, row_bias=rb, rows_per_sample=128, splitk=True), ref + res16.float() + rb.repeat_interleave(128, 0), what="split-K bf16 residual + row bias")
    o1, o2 = ops.gemm(a, w, bias=bias, splitk=True), ops.gemm(a, w, bias=bias, splitk=True)
    assert torch.equal(o1, o2), "split-K reduction must be deterministic"


def test_gemm_geglu():
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, Cc = 384, 320
    x = torch.randn(M, Cc, generator=g).to(bf).cuda()
    w = (torch.randn(8 * Cc, Cc, generator=g) / math.sqrt(Cc)).to(bf)
    b = torch.randn(8 * Cc, generator=g) * 0.1
    wi, bi = ops.pack_geglu_weight(w, b)
    got = ops.gemm(x, wi.cuda(), bias=bi.cuda(), geglu=True)
    wt, bt = ops.pack_geglu_weight_tiled(w.cuda(), b.cuda())
    assert torch.equal(got, ops.gemm(x, wt, bias=bt, geglu=True)), "tiled and plain GEGLU weights must give identical results"
    proj = x.float().cpu() @ w.float().t() + b
    h, gate = proj.chunk(2, -1)
    check(got, h * F.gelu(gate), what="geglu")


@pytest.mark.parametrize("M,N,K", [(4096, 1280, 1280), (16384, 640, 640), (4000, 1920, 640), (8192, 320, 2560), (2048, 3840, 1280)])
def test_gemm_cta_pair(M, N, K):
    """Tiled weights, no residual, enough rows: the GEMM runs as CTA pairs (tcgen05 cta_group::2, each CTA stages half of the weight tile)."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(M % 997 + N + K)
    a = torch.randn(M, K, generator=g).to(bf).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(bf).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref = a.float() @ w.float().t() + b
    for swz in (True, False):
        wt = ops.tile_weight(w, swizzle=swz)
        got = ops.gemm(a, wt, bias=b)
        check(got, ref, what=f"pair gemm swizzled={swz}")
        assert torch.equal(got, ops.gemm(a, w, bias=b)), "pair (tiled weights) and single-CTA (plain weights) paths must agree bit for bit"
    check(ops.gemm(a, ops.tile_weight(w), bias=b, out_f32=True), ref, tol=2e-3, what="pair gemm fp32 out")


def test_gemm_cta_pair_geglu():
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(77)
    M, Cc = 4096, 1280
    x = torch.randn(M, Cc, generator=g).to(bf).cuda()
    w = (torch.randn(8 * Cc, Cc, generator=g) / math.sqrt(Cc)).to(bf).cuda()
    b = (torch.randn(8 * Cc, generator=g) * 0.1).cuda()
    wt, bt = ops.pack_geglu_weight_tiled(w, b)
    got = ops.gemm(x, wt, bias=bt, geglu=True)
    proj = x.float() @ w.float().t() + b
    h, gate = proj.chunk(2, -1)
    check(got, h * F.gelu(gate), what="geglu pair")


def test_gemm_batched():
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(4)
    a = torch.randn(3, 256, 512, generator=g).to(bf).cuda()
    w = (torch.randn(3, 192, 512, generator=g) / 22).to(bf).cuda()
    check(ops.gemm(a, w), torch.bmm(a.float(), w.float().transpose(1, 2)), what="batched")


def _conv_ref(x_nhwc, w_oihw, bias=None, stride=1, upsample=False):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    if upsample:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    y = F.conv2d(x, w_oihw.float(), bias, stride=stride, padding=w_oihw.shape[-1] // 2)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 1280), (3, 16, 16, 1280, 640), (5, 8, 8, 1280, 1280),
                                            (1, 64, 64, 8, 320), (2, 64, 64, 320, 4), (1, 24, 40, 64, 128), (2, 128, 128, 128, 128)])
def test_conv3x3(N, H, W, Cin, Cout):
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(N * H + Cin + Cout)
    x = torch.randn(N, H, W, Cin, generator=g).to(bf)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)).to(bf)
    b = torch.randn(Cout, generator=g)
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight(w).cuda(), Cout, bias=b.cuda(), out_f32=(Cout == 4))
    check(got, _conv_ref(x, w, b), what=f"conv3x3 {N}x{H}x{W} {Cin}->{Cout}")
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w.cuda()), Cout, bias=b.cuda(), out_f32=(Cout == 4))
    check(got, _conv_ref(x, w, b), what=f"conv3x3 tiled-W {N}x{H}x{W} {Cin}->{Cout}")
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w.cuda()), Cout, bias=b.cuda(), out_f32=(Cout == 4), splitk=True)
    check(got, _conv_ref(x, w, b), what=f"conv3x3 split-K {N}x{H}x{W} {Cin}->{Cout}")


def test_conv_stride2_upsample_concat_epilogue():
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 32, 32, 320, generator=g).to(bf)
    w = (torch.randn(320, 320, 3, 3, generator=g) / math.sqrt(9 * 320)).to(bf)
    b = torch.randn(320, generator=g)
    wp = ops.pack_conv_weight(w).cuda()
    check(ops.conv2d(x.cuda(), wp, 320, stride=2, bias=b.cuda()), _conv_ref(x, w, b, stride=2), what="stride 2")
    check(ops.conv2d(x.cuda(), wp, 320, upsample=True, bias=b.cuda()), _conv_ref(x, w, b, upsample=True), what="upsample fold")
    # two-source concat + time-embedding row bias + residual
    x1 = torch.randn(2, 32, 32, 640, generator=g).to(bf)
    w2 = (torch.randn(320, 960, 3, 3, generator=g) / math.sqrt(9 * 960)).to(bf)
    rb = torch.randn(2, 320, generator=g)
    res = torch.randn(2, 32, 32, 320, generator=g).to(bf)
    ref = _conv_ref(torch.cat([x, x1], -1), w2, b) + rb[:, None, None, :] + res.float()
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight(w2).cuda(), 320, x1=x1.cuda(), bias=b.cuda(), row_bias=rb.cuda(), residual=res.cuda())
    check(got, ref, what="concat+temb+residual")
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w2.cuda()), 320, x1=x1.cuda(), bias=b.cuda(), row_bias=rb.cuda(), residual=res.cuda())
    check(got, ref, what="concat+temb+residual tiled-W")
    check(ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w.cuda()), 320, stride=2, bias=b.cuda()), _conv_ref(x, w, b, stride=2), what="stride 2 tiled-W")
    check(ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w.cuda()), 320, upsample=True, bias=b.cuda()), _conv_ref(x, w, b, upsample=True), what="upsample tiled-W")


@pytest.mark.parametrize("N,H,W,C0,C1,Cout", [(8, 64, 64, 320, 0, 320), (16, 32, 32, 640, 0, 640), (16, 16, 16, 1280, 0, 1280), (8, 64, 64, 640, 320, 320),
                                                (6, 32, 32, 1280, 640, 640), (9, 64, 64, 128, 0, 128),
                                                (5, 34, 32, 192, 0, 160), (3, 33, 64, 128, 0, 320),    # ragged last row block (128-row halo tiles)
                                                (8, 16, 16, 1280, 0, 1280), (7, 16, 16, 1280, 1280, 1280)])   # 128-row tiles run as CTA pairs (cta_group::2)
def test_conv3x3_halo_main_loop(N, H, W, C0, C1, Cout):
    """Problems big enough for the 256-row tiles: the 3x3 main loop loads one tall box per (horizontal tap, channel chunk) and the
    three vertical taps read it at shifted rows.  Checked against torch, against a batch of one (bit-exact) and with the epilogue extras."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(N + H + C0 + C1)
    x = torch.randn(N, H, W, C0, generator=g).to(bf)
    x1 = torch.randn(N, H, W, C1, generator=g).to(bf) if C1 else None
    w = (torch.randn(Cout, C0 + C1, 3, 3, generator=g) / math.sqrt(9 * (C0 + C1))).to(bf)
    b = torch.randn(Cout, generator=g)
    rb = torch.randn(N, Cout, generator=g)
    res = torch.randn(N, H, W, Cout, generator=g).to(bf)
    wt = ops.pack_conv_weight_tiled(w.cuda())
    full = torch.cat([x, x1], -1) if C1 else x
    ref = _conv_ref(full, w, b)
    got = ops.conv2d(x.cuda(), wt, Cout, x1=x1.cuda() if C1 else None, bias=b.cuda())
    check(got, ref, what=f"halo conv {N}x{H}x{W} {C0}+{C1}->{Cout}")
    got2 = ops.conv2d(x.cuda(), wt, Cout, x1=x1.cuda() if C1 else None, bias=b.cuda(), row_bias=rb.cuda(), residual=res.cuda())
    check(got2, ref + rb[:, None, None, :] + res.float(), what="halo conv + temb + residual")
    f32 = ops.conv2d(x.cuda(), wt, Cout, x1=x1.cuda() if C1 else None, bias=b.cuda(), out_f32=True)
    check(f32, ref, what="halo conv fp32 out")
    one = ops.conv2d(x[:1].cuda(), wt, Cout, x1=x1[:1].cuda() if C1 else None, bias=b.cuda())   # small batch: one-box-per-tap loop
    check(one, ref[:1], what="batch of one")


def test_conv_lowres_splitk_is_batch_independent():
    """The <= 8x8 layers split K four ways by a rule that only looks at the per-image geometry: a batch of 5 == five batches of 1,
    bit for bit, also when the partial-sum workspace only holds one tile of images at a time (chunked launches)."""
    from gm_diffusion_b200 import _lib as L
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(12)
    for hw, cin in ((8, 1280), (4, 1280), (8, 2560)):
        x = torch.randn(5, hw, hw, cin, generator=g).to(bf).cuda()
        w = (torch.randn(1280, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(bf)
        b = torch.randn(1280, generator=g).cuda()
        rb = torch.randn(5, 1280, generator=g).cuda()
        res = torch.randn(5, hw, hw, 1280, generator=g).to(bf).cuda()
        wt = ops.pack_conv_weight_tiled(w.cuda())
        L.lib().gmd_reset_launch_count()
        full = ops.conv2d(x, wt, 1280, bias=b, row_bias=rb, residual=res)
        assert L.lib().gmd_launch_count() == 2, "expected the split main loop + the finalize kernel"
        ref = _conv_ref(x.cpu(), w, b.cpu()) + rb.cpu()[:, None, None, :] + res.float().cpu()
        check(full, ref, what=f"low-res split-K conv {hw}x{hw} K={9 * cin}")
        for i in range(5):
            one = ops.conv2d(x[i:i + 1].contiguous(), wt, 1280, bias=b, row_bias=rb[i:i + 1].contiguous(), residual=res[i:i + 1].contiguous())
            assert torch.equal(one, full[i:i + 1]), (hw, cin, i)
        # chunked: a workspace that only fits one tile of images
        big = ops._WS.get((0, 0))   # (device 0, scratch slot 0)
        try:
            per_img = 4 * hw * hw * 1280 * 4
            imgs_per_tile = 128 // (hw * hw)
            ops._WS[(0, 0)] = torch.empty(per_img * imgs_per_tile + 64, dtype=torch.uint8, device="cuda")
            L.lib().gmd_reset_launch_count()
            chunked = ops.conv2d(x, wt, 1280, bias=b, row_bias=rb, residual=res)
            assert L.lib().gmd_launch_count() == 2 * -(-5 // imgs_per_tile)
        finally:
            ops._WS[(0, 0)] = big
        assert torch.equal(chunked, full), (hw, cin)


@pytest.mark.parametrize("N,H,W,C,Cout", [(2, 64, 64, 128, 128), (1, 32, 48, 256, 256), (1, 16, 16, 512, 512)])
def test_conv_stride2_pad_end(N, H, W, C, Cout):
    """AutoencoderKL encoder downsample: F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv with padding 0 (diffusers Downsample2D(padding=0))."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(N, H, W, C, generator=g).to(bf)
    w = (torch.randn(Cout, C, 3, 3, generator=g) / math.sqrt(9 * C)).to(bf)
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), b, stride=2, padding=0).permute(0, 2, 3, 1)
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight_tiled(w.cuda()), Cout, stride=2, pad_end=True, bias=b.cuda())
    assert got.shape == (N, H // 2, W // 2, Cout)
    check(got, ref, what="stride 2, bottom/right pad")


@pytest.mark.parametrize("B,H,Nq,Nk,d", [(2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160), (3, 8, 64, 64, 160),
                                         (2, 8, 4096, 77, 40), (2, 8, 1024, 77, 80), (2, 8, 256, 77, 160), (1, 8, 200, 130, 40),
                                         (1, 8, 256, 64, 80), (1, 8, 300, 130, 80), (2, 8, 512, 200, 80), (1, 8, 1000, 448, 80), (1, 8, 256, 128, 80),
                                         (1, 8, 300, 170, 40), (2, 8, 640, 1000, 40), (1, 8, 129, 161, 40)])   # ragged key counts through both key halves of a tile
def test_attention(B, H, Nq, Nk, d):
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(Nq + Nk + d)
    Cc = H * d
    if Nq == Nk:
        qkv = torch.randn(B, Nq, 3 * Cc, generator=g).to(bf).cuda()
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
    else:
        q = torch.randn(B, Nq, Cc, generator=g).to(bf).cuda()
        kv = torch.randn(B, Nk, 2 * Cc, generator=g).to(bf).cuda()
        k, v = kv[..., :Cc], kv[..., Cc:]
    got = ops.attention(q, k, v, H)
    qh, kh, vh = (t.float().reshape(B, -1, H, d).transpose(1, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh
    check(got, ref.transpose(1, 2).reshape(B, Nq, Cc), tol=1e-2, what=f"attn B{B} H{H} {Nq}x{Nk} d{d}")


@pytest.mark.parametrize("B,H,Nq,Nk,d", [(1, 8, 200, 77, 40), (3, 8, 384, 65, 40), (2, 8, 128, 80, 80), (1, 8, 1000, 77, 80), (2, 8, 16384, 77, 40),
                                         (5, 8, 4096, 77, 40), (1, 8, 77, 77, 40)])
def test_attention_text_cross(B, H, Nq, Nk, d):
    """The persistent text cross-attention kernel (64 < Nk <= 80, d = 40 / 80): ragged query counts, odd tile counts per CTA, both
    ends of the key range, and batch counts that leave partial waves."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(7 * Nq + Nk + d)
    Cc = H * d
    q = torch.randn(B, Nq, Cc, generator=g).to(bf).cuda()
    kv = torch.randn(B, Nk, 2 * Cc, generator=g).to(bf).cuda()
    k, v = kv[..., :Cc], kv[..., Cc:]
    got = ops.attention(q, k, v, H)
    qh, kh, vh = (t.float().reshape(B, -1, H, d).transpose(1, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh
    check(got, ref.transpose(1, 2).reshape(B, Nq, Cc), tol=1e-2, what=f"text cross-attn B{B} H{H} {Nq}x{Nk} d{d}")


@pytest.mark.parametrize("d,N,ramp", [(40, 512, False), (80, 512, False), (40, 1024, True), (80, 768, True)])
def test_attention_large_logits(d, N, ramp):
    """Peaky softmax: the running row maximum leaves the lazy window in later key tiles (the kernel notices from the tile's row sum and
    re-reads the true maximum from S).  `ramp` scales the keys up along the sequence, so that almost every tile raises the maximum and some
    exponents against the stale maximum exceed 128 (+inf in the discarded pass)."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(2 + d + N)
    B, H = 1, 8
    q = (torch.randn(B, N, H * d, generator=g) * 6).to(bf).cuda()
    k = torch.randn(B, N, H * d, generator=g) * 6
    if ramp:
        k = k * torch.linspace(0.05, 2.5, N).view(1, N, 1)
    k = k.to(bf).cuda()
    v = torch.randn(B, N, H * d, generator=g).to(bf).cuda()
    got = ops.attention(q, k, v, H)
    assert torch.isfinite(got.float()).all()
    qh, kh, vh = (t.float().reshape(B, -1, H, d).transpose(1, 2) for t in (q, k, v))
    ref = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, -1) @ vh
    check(got, ref.transpose(1, 2).reshape(B, N, H * d), tol=2e-2, what="peaky softmax (running-max rescale path)")


@pytest.mark.parametrize("case", ["conv3", "conv3_f32out_temb", "conv3_2src_res", "conv_up", "conv_s2", "gemm_res"])
def test_groupnorm_statistics_from_the_producer_epilogue(case):
    """north_star (b): the convolution / GEMM epilogue emits the GroupNorm statistics of its output; gmd_groupnorm_apply then
    normalises in one pass.  Checked: the statistics against sums over the stored output, the normalised result against torch's
    GroupNorm of that output, a two-source consumer whose 32 groups straddle the sources, bit-reproducibility and independence
    of the batch size."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(len(case))
    N, H, C = 3, 32, 320
    x = torch.randn(N, H, H, C, generator=g).to(bf).cuda()
    x1 = torch.randn(N, H, H, 640, generator=g).to(bf).cuda()
    mk = lambda co, ci: ops.pack_conv_weight_tiled((torch.randn(co, ci, 3, 3, generator=g) / math.sqrt(9 * ci)).cuda())
    b = torch.randn(640, generator=g).cuda()
    if case == "conv3":
        run = lambda xx, n: ops.conv2d(xx, w, 640, bias=b, gn_stats=True); w = mk(640, C)
    elif case == "conv3_f32out_temb":
        rb = torch.randn(N, 640, generator=g).cuda(); w = mk(640, C)
        run = lambda xx, n: ops.conv2d(xx, w, 640, bias=b, row_bias=rb[:n], out_f32=True, gn_stats=True)
    elif case == "conv3_2src_res":
        w = mk(640, C + 640); res = torch.randn(N, H, H, 640, generator=g).to(bf).cuda()
        run = lambda xx, n: ops.conv2d(xx, w, 640, x1=x1[:n], bias=b, residual=res[:n], gn_stats=True)
    elif case == "conv_up":
        w = mk(640, C)
        run = lambda xx, n: ops.conv2d(xx, w, 640, upsample=True, bias=b, gn_stats=True)
    elif case == "conv_s2":
        w = mk(640, C)
        run = lambda xx, n: ops.conv2d(xx, w, 640, stride=2, bias=b, gn_stats=True)
    else:
        wl = ops.tile_weight((torch.randn(640, C, generator=g) / math.sqrt(C)).cuda()); res = torch.randn(N * H * H, 640, generator=g).to(bf).cuda()
        run = lambda xx, n: ops.gemm(xx.reshape(-1, C), wl, bias=b, residual=res[: n * H * H], gn_rows_per_sample=H * H)
    y, sums = run(x, N)
    assert sums is not None and sums.shape == (N, 320, 2), "this shape must take the epilogue-statistics path"
    y3 = y.float().reshape(N, -1, 640)
    want = torch.stack([y3.reshape(N, -1, 320, 2).sum((1, 3)), (y3 ** 2).reshape(N, -1, 320, 2).sum((1, 3))], -1)
    assert sums.dtype == torch.int64                     # fixed point, units of 2^-24: integer atomics are order-independent
    fsums = sums.double().cpu() / 2 ** 24
    err = (fsums - want.double().cpu()).abs().max() / want.abs().max()
    assert float(err) < 2e-3, f"{case}: epilogue statistics vs sums over the stored output: {float(err):.2e}"   # (formed before the bf16 rounding of the output)
    ga, be = torch.randn(640, generator=g).cuda(), torch.randn(640, generator=g).cuda()
    got = ops.groupnorm_silu(y.reshape(N, -1, 640), ga, be, eps=1e-5, sums=sums)
    ref = F.silu(F.group_norm(y3.permute(0, 2, 1), 32, ga, be, 1e-5)).permute(0, 2, 1)
    check(got, ref, what=f"{case}: groupnorm from epilogue statistics")
    y_, sums_ = run(x, N)
    assert torch.equal(sums, sums_) and torch.equal(y, y_), "bit-reproducible"
    y1, s1 = run(x[:1], 1)
    assert torch.equal(s1, sums[:1]) and torch.equal(y1.reshape(1, -1, 640), y.reshape(N, -1, 640)[:1]), "independent of the batch size"
    # consumer of a concat whose groups straddle the two sources (640 + 320 channels: groups of 30)
    n_px = y3.shape[1]
    z = torch.randn(N, n_px, 320, generator=g).to(bf).cuda()
    zs = (torch.stack([z.double().reshape(N, n_px, 160, 2).sum((1, 3)), (z.double() ** 2).reshape(N, n_px, 160, 2).sum((1, 3))], -1) * 2 ** 24).round().long().contiguous()
    ga2, be2 = torch.randn(960, generator=g).cuda(), torch.randn(960, generator=g).cuda()
    if y.dtype == bf:
        got2 = ops.groupnorm_silu(y.reshape(N, n_px, 640), ga2, be2, x1=z, sums=sums, sums1=zs)
        ref2 = F.silu(F.group_norm(torch.cat([y3, z.float()], -1).permute(0, 2, 1), 32, ga2, be2, 1e-5)).permute(0, 2, 1)
        check(got2, ref2, what=f"{case}: two-source groupnorm from epilogue statistics")


@pytest.mark.parametrize("N,HW,C0,C1", [(2, 4096, 320, 0), (2, 1024, 1280, 640), (3, 64, 1280, 1280), (1, 256, 640, 320), (2, 4096, 128, 0), (1, 100, 512, 0),
                                        (1, 16384, 512, 0), (1, 65536, 128, 0), (2, 4096, 640, 320)])  # 65536: too big for the one-pass cluster kernel -> two-pass
def test_groupnorm(N, HW, C0, C1):
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(HW + C0 + C1)
    x = (torch.randn(N, HW, C0, generator=g) * 2 + 0.5).to(bf)
    x1 = (torch.randn(N, HW, C1, generator=g) - 1).to(bf) if C1 else None
    Cc = C0 + C1
    ga, be = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
    got = ops.groupnorm_silu(x.cuda(), ga.cuda(), be.cuda(), x1=x1.cuda() if C1 else None, eps=1e-5)
    full = torch.cat([x, x1], -1) if C1 else x
    ref = F.silu(F.group_norm(full.float().permute(0, 2, 1), 32, ga, be, 1e-5)).permute(0, 2, 1)
    check(got, ref, what=f"groupnorm+silu {N}x{HW}x{C0}+{C1}")
    got = ops.groupnorm_silu(x.cuda(), ga[:C0].cuda(), be[:C0].cuda(), eps=1e-6, silu=False)
    check(got, F.group_norm(x.float().permute(0, 2, 1), 32, ga[:C0], be[:C0], 1e-6).permute(0, 2, 1), what="groupnorm no silu")
    again = ops.groupnorm_silu(x.cuda(), ga[:C0].cuda(), be[:C0].cuda(), eps=1e-6, silu=False)
    assert torch.equal(got, again), "groupnorm must be bit-reproducible"
    one = ops.groupnorm_silu(x[:1].cuda(), ga[:C0].cuda(), be[:C0].cuda(), eps=1e-6, silu=False)
    assert torch.equal(got[:1], one), "groupnorm of a sample must not depend on the batch size"
    x32 = torch.randn(N, HW, C0, generator=g) * 2 + 0.5
    got = ops.groupnorm_silu(x32.cuda(), ga[:C0].cuda(), be[:C0].cuda(), eps=1e-5)
    check(got, F.silu(F.group_norm(x32.permute(0, 2, 1), 32, ga[:C0], be[:C0], 1e-5)).permute(0, 2, 1), what="groupnorm fp32-in")


@pytest.mark.parametrize("M,Cc", [(8192, 320), (2048, 640), (513, 1280), (7, 512), (1001, 320), (5, 640), (1, 320)])
def test_layernorm(M, Cc):
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(M + Cc)
    x = (torch.randn(M, Cc, generator=g) * 3 + 1).to(bf)
    ga, be = torch.randn(Cc, generator=g), torch.randn(Cc, generator=g)
    check(ops.layernorm(x.cuda(), ga.cuda(), be.cuda()), F.layer_norm(x.float(), (Cc,), ga, be, 1e-5), what=f"layernorm {M}x{Cc}")
    x32 = torch.randn(M, Cc, generator=g) * 3 + 1
    check(ops.layernorm(x32.cuda(), ga.cuda(), be.cuda()), F.layer_norm(x32, (Cc,), ga, be, 1e-5), what=f"layernorm fp32-in {M}x{Cc}")


@pytest.mark.parametrize("M,Cc,Nq", [(4096, 320, 960), (1024, 640, 640), (1000, 1280, 1280), (8192, 320, 320)])
def test_layernorm_folded_into_the_gemms_around_it(M, Cc, Nq):
    """gmd_b200.h "LayerNorm folded into the GEMMs on either side of it": the GEMM that writes the fp32 token stream also emits a bf16
    copy and exact per-row sums; the projection behind the LayerNorm multiplies the copy with W diag(gamma) and normalises in its
    epilogue.  Reference: fp32 torch of  LN(x_out) @ W^T  with x_out = a @ Wo^T + b + residual."""
    from gm_diffusion_b200 import ops
    g = torch.Generator().manual_seed(M + Cc + Nq)
    a = torch.randn(M, Cc, generator=g).to(bf)
    wo = (torch.randn(Cc, Cc, generator=g) / Cc ** 0.5).to(bf)
    bo = torch.randn(Cc, generator=g)
    res = torch.randn(M, Cc, generator=g) * 2 + 0.7            # a mean the normalisation has to remove
    ga, be = 1 + 0.3 * torch.randn(Cc, generator=g), 0.2 * torch.randn(Cc, generator=g)
    w = torch.randn(Nq, Cc, generator=g) / Cc ** 0.5
    x_ref = a.float() @ wo.float().t() + bo + res
    want = F.layer_norm(x_ref, (Cc,), ga, be, 1e-5) @ w.t()
    # producer
    x, xb, xs = ops.gemm(a.cuda(), ops.tile_weight(wo.cuda()), bias=bo.cuda(), residual=res.cuda(), out_f32=True, ln_out=True)
    check(x, x_ref, tol=3e-3, what="token stream")
    assert torch.equal(xb, x.to(bf)), "the bf16 copy is the rounded stream"
    s1 = xs[:, 0].double() / 2 ** 24
    s2 = xs[:, 1].double() / 2 ** 24
    assert torch.allclose(s1.cpu(), x.double().sum(1).cpu(), rtol=0, atol=1e-3) and torch.allclose(s2.cpu(), x.double().pow(2).sum(1).cpu(), rtol=1e-6, atol=1e-3)
    # consumer
    wg = (w * ga[None, :]).to(bf)
    got = ops.gemm(xb, ops.tile_weight(wg.cuda()), bias=(w @ be).cuda(), ln_in=(xs, wg.float().sum(1).cuda(), 1e-5))
    check(got, want, tol=8e-3, what=f"folded LayerNorm {M}x{Cc}->{Nq}")
    # and it agrees with the unfolded path (LayerNorm kernel, then the plain projection) to bf16 rounding
    plain = ops.gemm(ops.layernorm(x, ga.cuda(), be.cuda()), ops.tile_weight(w.to(bf).cuda()))
    assert rel_l2(got, plain) < 8e-3
    # statistics are accumulated: a second producer call into the same (not re-zeroed) buffer would double them, so every call gets a fresh one
    _, _, xs2 = ops.gemm(a.cuda(), ops.tile_weight(wo.cuda()), bias=bo.cuda(), residual=res.cuda(), out_f32=True, ln_out=True)
    assert torch.equal(xs, xs2), "row statistics are bit-reproducible"


def test_softmax_silu_temb():
    from gm_diffusion_b200 import ops
    from oracle.unet_oracle import timestep_embedding
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(300, 4096, generator=g) * 4).to(bf)
    check(ops.softmax_rows(x.cuda(), 0.044), torch.softmax(x.float() * 0.044, -1), what="softmax rows")
    check(ops.silu(x.cuda()), F.silu(x.float()), what="silu")
    for t in (981.0, 501.0, 1.0):
        check(ops.timestep_embedding(t, 3, 320, "cuda"), timestep_embedding(torch.full((3,), t), 320), what=f"temb {t}")
