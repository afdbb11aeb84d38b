"""Per-kernel parity on the B200: every C-ABI entry point against a plain PyTorch fp32 reference of the same
op on identical (bf16-rounded) inputs.  Tolerance: relative L2 <= 1e-2 for bf16 outputs,
tighter for fp32 outputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

LEAK = 0.2
GAIN = math.sqrt(2.0 / (1.0 + LEAK * LEAK))


def ops():
    from neuron_gan_b200 import ops as o
    return o


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return bf(torch.randn(*shape, device='cuda', generator=g) * scale)


def pn_ref(h):
    r = torch.rsqrt((h * h).mean(1, keepdim=True) + 1e-8)
    return h * r, r


CONV_SHAPES = [  # B, cin, cout, H, W
    (2, 16, 16, 16, 16), (1, 16, 16, 64, 64), (2, 16, 32, 32, 32), (2, 32, 16, 32, 32), (1, 32, 32, 128, 128),
    (2, 32, 64, 32, 32), (2, 64, 32, 32, 32), (2, 64, 64, 32, 32), (2, 64, 128, 16, 16), (2, 128, 64, 32, 32),
    (3, 128, 128, 16, 16), (1, 16, 16, 256, 256), (1, 16, 16, 40, 72),
    (9, 16, 16, 256, 256), (40, 32, 32, 64, 64), (48, 64, 64, 32, 32),     # several tiles per persistent CTA
]
# the tile plans of the benched configuration (BASELINE config 3: 512x512, 16 images per GPU; the critic batches
# [real; fake] -> 32) -- the launches bench.py's roofline line is measured on
BENCH_SHAPES = [(16, 16, 16, 512, 512), (32, 16, 16, 512, 512), (16, 32, 16, 256, 256), (32, 16, 32, 128, 128)]
CONV_SHAPES_ALL = CONV_SHAPES + BENCH_SHAPES


def test_layout_roundtrip():
    o = ops()
    x = rnd(2, 32, 8, 12)
    c8 = o.nchw_to_c8(x)
    assert c8.shape == (2, 4, 8, 12, 8)
    assert torch.equal(c8.float().permute(0, 1, 4, 2, 3).reshape(2, 32, 8, 12), x)
    assert torch.equal(o.c8_to_nchw(c8), x)


@pytest.mark.parametrize('B,cin,cout,H,W', CONV_SHAPES_ALL)
@pytest.mark.parametrize('use_bias', [False, True])
def test_conv3x3_fwd(B, cin, cout, H, W, use_bias):
    o = ops()
    x = rnd(B, cin, H, W, seed=1)
    w = rnd(cout, cin, 3, 3, seed=2, scale=GAIN / math.sqrt(cin * 9))
    bias = rnd(cout, seed=3, scale=0.1) if use_bias else None
    s = GAIN / math.sqrt(cin * 9)
    h = F.leaky_relu(F.conv2d(s * x, w, bias, padding=1), LEAK)
    y_ref, r_ref = pn_ref(h)
    w_fwd, _ = o.prep_conv_weight(w)
    y, r = o.conv3x3_fwd(o.nchw_to_c8(x), w_fwd, bias, s, LEAK, cout)
    torch.cuda.synchronize()
    assert rel(o.c8_to_nchw(y), y_ref) < 6e-3
    assert rel(r, r_ref[:, 0]) < 2e-3


@pytest.mark.parametrize('B,cin,cout,H,W', CONV_SHAPES_ALL)
def test_conv3x3_dgrad_and_pn(B, cin, cout, H, W):
    o = ops()
    ga = rnd(B, cout, H, W, seed=4)
    w = rnd(cout, cin, 3, 3, seed=5, scale=GAIN / math.sqrt(cin * 9))
    s = GAIN / math.sqrt(cin * 9)
    gx_ref = s * F.conv_transpose2d(ga, w, padding=1)
    _, w_dg = o.prep_conv_weight(w)
    gx = o.conv3x3_dgrad(o.nchw_to_c8(ga), w_dg, s, cin)
    assert rel(o.c8_to_nchw(gx), gx_ref) < 6e-3
    # fused PixelNorm/LeakyReLU backward of the producing layer
    h = rnd(B, cin, H, W, seed=6)
    y_prev, r_prev = pn_ref(F.leaky_relu(h, LEAK))
    y_prev = bf(y_prev)
    addin = rnd(B, cin, H, W, seed=7, scale=0.1)
    mask = torch.where(y_prev > 0, 1.0, LEAK)
    ga_prev_ref = mask * r_prev * (gx_ref - y_prev * (gx_ref * y_prev).mean(1, keepdim=True)) + addin
    ga_prev, gy = o.conv3x3_dgrad_pn(o.nchw_to_c8(ga), w_dg, s, LEAK, o.nchw_to_c8(y_prev),
                                     r_prev[:, 0].contiguous(), addin=o.nchw_to_c8(addin), want_gy=True)
    assert rel(o.c8_to_nchw(ga_prev), ga_prev_ref) < 8e-3
    assert rel(o.c8_to_nchw(gy), gx_ref) < 6e-3


@pytest.mark.parametrize('B,cin,cout,H,W', CONV_SHAPES[:11] + BENCH_SHAPES[:1] + BENCH_SHAPES[2:])
def test_conv3x3_double_backward(B, cin, cout, H, W):
    """conv3x3_dbl against autograd's double backward of conv -> LeakyReLU -> PixelNorm."""
    o = ops()
    s = GAIN / math.sqrt(cin * 9)
    w = rnd(cout, cin, 3, 3, seed=8, scale=GAIN / math.sqrt(cin * 9))
    x = rnd(B, cin, H, W, seed=9)
    a = (s * F.conv2d(x, w, padding=1)).detach().requires_grad_()
    h = F.leaky_relu(a, LEAK)
    r = torch.rsqrt((h * h).mean(1, keepdim=True) + 1e-8)
    y = h * r
    gy = rnd(B, cout, H, W, seed=10).requires_grad_()
    ga, = torch.autograd.grad(y, a, gy, create_graph=True)
    gx = s * F.conv_transpose2d(ga, w, padding=1)
    ghat_x = rnd(B, cin, H, W, seed=11)
    cot_gy, cot_a = torch.autograd.grad(gx, (gy, a), ghat_x)
    w_fwd, _ = o.prep_conv_weight(w)
    yb = bf(y.detach())
    ghat_y, ahat = o.conv3x3_dbl(o.nchw_to_c8(ghat_x), w_fwd, s, LEAK, o.nchw_to_c8(yb), r.detach()[:, 0].contiguous(),
                                 o.nchw_to_c8(gy.detach()))
    assert rel(o.c8_to_nchw(ghat_y), cot_gy) < 1.5e-2
    assert rel(o.c8_to_nchw(ahat), cot_a) < 2e-2


@pytest.mark.parametrize('B,cin,cout,H,W', CONV_SHAPES[:12] + CONV_SHAPES[13:] + BENCH_SHAPES)
def test_conv3x3_wgrad(B, cin, cout, H, W):
    o = ops()
    x = rnd(B, cin, H, W, seed=12)
    ga = rnd(B, cout, H, W, seed=13)
    s = 0.37
    dw_ref = s * torch.nn.grad.conv2d_weight(x, (cout, cin, 3, 3), ga, padding=1)
    dw = torch.zeros(cout, cin, 3, 3, device='cuda')
    xc, gc = o.nchw_to_c8(x), o.nchw_to_c8(ga)
    o.conv3x3_wgrad(xc, gc, s, dw)
    assert rel(dw, dw_ref) < 2e-3
    first = dw.clone()
    o.conv3x3_wgrad(xc, gc, s, dw)                                     # accumulates
    assert rel(dw, 2 * dw_ref) < 2e-3
    # overwrite mode needs no zeroing, and the reduction is deterministic: bit-identical from run to run
    dw2 = torch.full_like(dw, float('nan'))
    o.conv3x3_wgrad(xc, gc, s, dw2, accumulate=False)
    assert torch.equal(dw2, first)


@pytest.mark.parametrize('C,H,W', [(16, 8, 8), (32, 16, 24), (128, 16, 16)])
def test_resampling(C, H, W):
    o = ops()
    x = rnd(2, C, H, W, seed=14)
    up = o.c8_to_nchw(o.upsample2x(o.nchw_to_c8(x)))
    assert rel(up, F.interpolate(x, scale_factor=2, mode='bilinear')) < 4e-3
    pool = o.c8_to_nchw(o.avgpool2(o.nchw_to_c8(x)))
    assert rel(pool, F.avg_pool2d(x, 2)) < 4e-3
    # adjoint of the upsample fused with PixelNorm/LeakyReLU backward
    h = rnd(2, C, H, W, seed=15)
    y, r = pn_ref(F.leaky_relu(h, LEAK))
    y = bf(y)
    g_up = rnd(2, C, 2 * H, 2 * W, seed=16)
    xr = x.clone().requires_grad_()
    g_low, = torch.autograd.grad(F.interpolate(xr, scale_factor=2, mode='bilinear'), xr, g_up)
    extra_pre = rnd(2, H, W, seed=17)
    extra_w = rnd(C, seed=18)
    g_tot = g_low + extra_w.view(1, C, 1, 1) * extra_pre.unsqueeze(1)
    mask = torch.where(y > 0, 1.0, LEAK)
    ref = mask * r * (g_tot - y * (g_tot * y).mean(1, keepdim=True))
    ga = o.up2_bwd_pn_bwd(o.nchw_to_c8(g_up), o.nchw_to_c8(y), r[:, 0].contiguous(), extra_pre, extra_w)
    assert rel(o.c8_to_nchw(ga), ref) < 8e-3
    # pn_bwd with un-pooling
    g_pool = rnd(2, C, H // 2, W // 2, seed=19)
    g_full = F.interpolate(g_pool, scale_factor=2, mode='nearest') * 0.25
    addin = rnd(2, C, H, W, seed=20)
    ref = mask * r * (g_full - y * (g_full * y).mean(1, keepdim=True)) + addin
    ga, gy = o.pn_bwd(o.nchw_to_c8(g_pool), o.nchw_to_c8(y), r[:, 0].contiguous(), gscale=0.25, unpool=True,
                      addin=o.nchw_to_c8(addin), want_gy=True)
    assert rel(o.c8_to_nchw(ga), ref) < 8e-3
    assert rel(o.c8_to_nchw(gy), g_full) < 4e-3


def test_image_ops():
    o = ops()
    x = torch.randn(3, 16, 24, device='cuda')
    assert torch.allclose(o.pool_image(x), F.avg_pool2d(x.unsqueeze(1), 2)[:, 0], atol=1e-6)
    assert torch.allclose(o.up2_image(x), F.interpolate(x.unsqueeze(1), scale_factor=2, mode='bilinear')[:, 0], atol=1e-6)
    g = torch.randn(3, 32, 48, device='cuda')
    xr = x.clone().requires_grad_()
    ref, = torch.autograd.grad(F.interpolate(xr.unsqueeze(1), scale_factor=2, mode='bilinear'), xr, g.unsqueeze(1))
    assert torch.allclose(o.up2_image_bwd(g, 0.5), 0.5 * ref, atol=1e-5)
    assert torch.allclose(o.unpool_image(x, 0.25), 0.25 * F.interpolate(x.unsqueeze(1), scale_factor=2)[:, 0])
    y = torch.randn_like(x)
    assert torch.allclose(o.lerp(x, y, 0.3), x + 0.3 * (y - x), atol=1e-6)
    eps = torch.rand(3, device='cuda')
    assert torch.allclose(o.interp_images(x, y, eps), eps.view(3, 1, 1) * x + (1 - eps.view(3, 1, 1)) * y, atol=1e-6)
    assert torch.allclose(o.scale_rows(x, eps, 2.0), 2 * eps.view(3, 1, 1) * x, atol=1e-6)


@pytest.mark.parametrize('misalign', [False, True])
def test_image_ops_vector_path(misalign):
    """The four-floats-per-thread variants (rows >= 128 floats, 16-byte aligned) against torch, and the same shapes
    through the scalar kernels when the base pointer is only 4-byte aligned."""
    o = ops()

    def mk(*shape):
        n = 1
        for d in shape:
            n *= d
        flat = torch.randn(n + 1, device='cuda')
        return flat[1:].view(*shape) if misalign else flat[:n].view(*shape)

    x, y = mk(3, 64, 256), mk(3, 64, 256)
    assert (x.data_ptr() % 16 != 0) == misalign
    assert torch.allclose(o.pool_image(x), F.avg_pool2d(x.unsqueeze(1), 2)[:, 0], atol=1e-6)
    assert torch.equal(o.unpool_image(x, 0.25), 0.25 * F.interpolate(x.unsqueeze(1), scale_factor=2)[:, 0])
    eps = torch.rand(3, device='cuda')
    e = eps.view(3, 1, 1)
    assert torch.allclose(o.interp_images(x, y, eps), e * x + (1 - e) * y, atol=1e-6)
    assert torch.allclose(o.scale_rows(x, eps, 2.0), 2 * e * x, atol=1e-6)
    assert torch.allclose(o.lerp(x, y, 0.3), x + 0.3 * (y - x), atol=1e-6)
    assert torch.allclose(o.up2_image(x), F.interpolate(x.unsqueeze(1), scale_factor=2, mode='bilinear')[:, 0], atol=2e-6)
    g = mk(3, 128, 512)
    xr = x.clone().requires_grad_()
    ref, = torch.autograd.grad(F.interpolate(xr.unsqueeze(1), scale_factor=2, mode='bilinear'), xr, g.unsqueeze(1))
    assert torch.allclose(o.up2_image_bwd(g, 0.5), 0.5 * ref, atol=1e-5)
    # whatever the alignment, the same values (vector against scalar kernels)
    xa = x.clone()
    assert torch.allclose(o.pool_image(xa), o.pool_image(x), rtol=0, atol=1e-7)
    assert torch.allclose(o.interp_images(xa, y.clone(), eps), o.interp_images(x, y, eps), rtol=0, atol=1e-6)


@pytest.mark.parametrize('C', [16, 128])
def test_fromim_toim(C):
    o = ops()
    B, H, W = 2, 16, 16
    xp = torch.randn(B, H, W, device='cuda')
    w, b = torch.randn(C, device='cuda'), torch.randn(C, device='cuda')
    f_ref = w.view(1, C, 1, 1) * xp.unsqueeze(1) + b.view(1, C, 1, 1)
    assert rel(o.c8_to_nchw(o.fromim_fwd(xp, w, b)), f_ref) < 4e-3
    y_end = rnd(B, C, H, W, seed=21)
    fade = o.d_fade_fwd(o.nchw_to_c8(y_end), xp, w, b, 0.3)
    assert rel(o.c8_to_nchw(fade), f_ref + 0.3 * (y_end - f_ref)) < 4e-3
    g = rnd(B, C, H, W, seed=22)
    gw, gb = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    g_img = torch.empty(B, H, W, device='cuda')
    o.fromim_bwd(o.nchw_to_c8(g), xp, w, gw, gb, gscale=0.5, g_img=g_img)
    assert rel(gw, 0.5 * (g * xp.unsqueeze(1)).sum((0, 2, 3))) < 1e-4
    assert rel(gb, 0.5 * g.sum((0, 2, 3))) < 1e-4
    assert rel(g_img, 0.5 * (g * w.view(1, C, 1, 1)).sum(1)) < 1e-4
    # double backward
    ghat = torch.randn(B, H, W, device='cuda')
    what = torch.zeros(C, device='cuda')
    out = o.fromim_dbl(ghat, o.nchw_to_c8(g), w, what, in_scale=2.0, gscale=0.5)
    assert rel(o.c8_to_nchw(out), 2.0 * w.view(1, C, 1, 1) * ghat.unsqueeze(1)) < 4e-3
    assert rel(what, (2.0 * ghat.unsqueeze(1) * 0.5 * g).sum((0, 2, 3))) < 1e-4
    # ToImage
    h = rnd(B, C, H, W, seed=23)
    y, r = pn_ref(F.leaky_relu(h, LEAK))
    y = bf(y)
    wt = torch.randn(C, device='cuda') * 0.2
    img_ref = torch.tanh((y * wt.view(1, C, 1, 1)).sum(1))
    img = o.toim_fwd(o.nchw_to_c8(y), wt)
    assert torch.allclose(img, img_ref, atol=2e-5)
    g_img = torch.randn(B, H, W, device='cuda')
    gpre_ref = 0.7 * g_img * (1 - img_ref ** 2)
    gy = wt.view(1, C, 1, 1) * gpre_ref.unsqueeze(1)
    mask = torch.where(y > 0, 1.0, LEAK)
    ga_ref = mask * r * (gy - y * (gy * y).mean(1, keepdim=True))
    gw = torch.zeros(C, device='cuda')
    ga, gpre = o.toim_bwd(g_img, img, o.nchw_to_c8(y), r[:, 0].contiguous(), wt, gw, gscale=0.7, want_gpre=True)
    assert rel(gpre, gpre_ref) < 1e-4
    assert rel(o.c8_to_nchw(ga), ga_ref) < 8e-3
    assert rel(gw, (gpre_ref.unsqueeze(1) * y).sum((0, 2, 3))) < 1e-4
    # overwrite mode (no zeroing) gives the same bits as accumulation into zeros, run after run
    gw_o = torch.full((C,), float('nan'), device='cuda')
    o.toim_bwd(g_img, img, o.nchw_to_c8(y), r[:, 0].contiguous(), wt, gw_o, gscale=0.7, grad_accumulate=False)
    assert torch.equal(gw_o, gw)
    gw_f, gb_f = torch.full((C,), float('nan'), device='cuda'), torch.full((C,), float('nan'), device='cuda')
    o.fromim_bwd(o.nchw_to_c8(g), xp, w, gw_f, gb_f, gscale=0.5, grad_accumulate=False)
    assert rel(gw_f, 0.5 * (g * xp.unsqueeze(1)).sum((0, 2, 3))) < 1e-4 and rel(gb_f, 0.5 * g.sum((0, 2, 3))) < 1e-4
    what_o = torch.full((C,), float('nan'), device='cuda')
    o.fromim_dbl(ghat, o.nchw_to_c8(g), w, what_o, in_scale=2.0, gscale=0.5, grad_accumulate=False)
    assert torch.equal(what_o, what)
    # the fade kernel without a bias (the double backward's cotangent blend)
    fade0 = o.d_fade_fwd(o.nchw_to_c8(y_end), xp, w, None, 0.3)
    f0 = w.view(1, C, 1, 1) * xp.unsqueeze(1)
    assert rel(o.c8_to_nchw(fade0), f0 + 0.3 * (y_end - f0)) < 4e-3


@pytest.mark.parametrize('B', [1, 5])
def test_head(B):
    o = ops()
    C, S = 128, 16
    h = rnd(B, C, S, S, seed=24)
    y, r = pn_ref(F.leaky_relu(h, LEAK))
    y = bf(y)
    w = torch.randn(1, C, S, S, device='cuda') * 0.01
    bias = torch.randn(1, device='cuda')
    s = GAIN / math.sqrt(C * S * S)
    ref = F.conv2d(s * y, w, bias).flatten()
    score = o.head_fwd(o.nchw_to_c8(y), w, bias, s)
    assert torch.allclose(score, ref, rtol=1e-4, atol=1e-5)
    gout = torch.randn(B, device='cuda')
    gy_ref = s * w * gout.view(B, 1, 1, 1)
    mask = torch.where(y > 0, 1.0, LEAK)
    ga_ref = mask * r * (gy_ref - y * (gy_ref * y).mean(1, keepdim=True))
    ga, gy = o.head_bwd_pn(gout, w, s, o.nchw_to_c8(y), r[:, 0].contiguous(), want_gy=True)
    assert rel(o.c8_to_nchw(ga), ga_ref) < 8e-3
    assert rel(o.c8_to_nchw(gy), gy_ref) < 4e-3
    gw, ghb = torch.zeros_like(w), torch.full((1,), 0.5, device='cuda')
    o.head_wgrad(o.nchw_to_c8(y), gout, s, gw, ghb)          # accumulates: gw += ..., bias gradient += sum(gout)
    o.head_wgrad(o.nchw_to_c8(y), gout, s, gw)
    assert rel(gw, 2 * s * (y * gout.view(B, 1, 1, 1)).sum(0, keepdim=True)) < 1e-4
    assert torch.allclose(ghb, 0.5 + gout.sum(), atol=1e-5)
    gb = torch.zeros(C, device='cuda')
    o.bias_grad(o.nchw_to_c8(y), gb)
    assert rel(gb, y.sum((0, 2, 3))) < 1e-4
    # overwrite mode, bit-reproducible (no atomics anywhere in the parameter-gradient kernels)
    gw2, gb2 = torch.full_like(gw, float('nan')), torch.full((C,), float('nan'), device='cuda')
    o.head_wgrad(o.nchw_to_c8(y), gout, s, gw2, accumulate=False)
    o.bias_grad(o.nchw_to_c8(y), gb2, accumulate=False)
    assert rel(gw2, s * (y * gout.view(B, 1, 1, 1)).sum(0, keepdim=True)) < 1e-4 and torch.equal(gb2, gb)


@pytest.mark.parametrize('B', [3, 16, 40])
def test_linear(B):
    o = ops()
    C, S, K = 128, 16, 512
    w = rnd(C * S * S, K, seed=25, scale=GAIN / math.sqrt(K))
    z = torch.randn(B, K, device='cuda')
    z = z / z.norm(dim=1, keepdim=True)
    s = GAIN / math.sqrt(K)
    h = F.leaky_relu(F.linear(s * z, w), LEAK).unflatten(1, (C, S, S))
    y_ref, r_ref = pn_ref(h)
    wb = o.prep_linear_weight(w, C, S)
    assert torch.equal(o.linear_image_to_matrix(wb, K, C, S).float(), w.to(torch.bfloat16).float())
    y, r = o.linear_fwd(z, wb, s, LEAK, C, S)
    assert rel(o.c8_to_nchw(y), y_ref) < 5e-3
    assert rel(r, r_ref[:, 0]) < 1e-3
    ga = rnd(B, C, S, S, seed=26)
    dw = torch.zeros_like(w)
    o.linear_wgrad(o.nchw_to_c8(ga), z, s, dw)
    assert rel(dw, s * ga.flatten(1).t() @ z) < 1e-4


@pytest.mark.parametrize('B,n_seg', [(3, 1), (16, 1), (64, 1), (16, 8), (5, 2), (64, 4), (70, 1)])
def test_adam_linear_factored(B, n_seg):
    """ops.adam_linear_factored (gradient of the Linear weight formed from its factors inside the Adam pass, tensor
    cores, z = hi + lo) against linear_wgrad + torch.optim.Adam on the materialised gradient; samples in per-rank
    segments of an all-gathered buffer (n_seg > 1) give the same result as one contiguous batch."""
    o = ops()
    C, S, K = 128, 16, 512
    torch.manual_seed(B * 10 + n_seg)
    p0 = torch.randn(C * S * S, K, device='cuda') * 0.05
    Bt = B * n_seg
    ga = rnd(Bt, C, S, S, seed=27, scale=0.01)
    z = torch.randn(Bt, K, device='cuda')
    z = z / z.norm(dim=1, keepdim=True)
    s = GAIN / math.sqrt(K) / n_seg
    g_ref = torch.zeros_like(p0)
    o.linear_wgrad(o.nchw_to_c8(ga), z, s, g_ref, accumulate=False)
    assert rel(g_ref, s * ga.flatten(1).t() @ z) < 1e-4
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.5, 0.999))
    own = torch.nn.Parameter(p0.clone())
    opt_own = torch.optim.Adam([own], lr=1e-3, betas=(0.5, 0.999))
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    img = torch.zeros(p.numel(), dtype=torch.bfloat16, device='cuda')
    g_out = torch.empty_like(p0)
    # the factors as an all-gathered buffer: n_seg segments [ga rows | z rows], with padding between the segments
    ga_c8 = o.nchw_to_c8(ga)
    n_ga, n_z = B * C * S * S, B * K
    per = n_ga * 2 + n_z * 4 + 64
    buf = torch.zeros((n_seg, per), dtype=torch.uint8, device='cuda')
    buf_ga = buf[:, :n_ga * 2].view(torch.bfloat16)
    buf_z = buf[:, n_ga * 2:n_ga * 2 + n_z * 4].view(torch.float32)
    for r in range(n_seg):
        buf_ga[r].copy_(ga_c8[r * B:(r + 1) * B].reshape(-1))
        buf_z[r].copy_(z[r * B:(r + 1) * B].reshape(-1))
    for step in range(1, 3):
        ref.grad = g_ref.clone()
        opt.step()
        o.adam_linear_factored(p, m, v, img, buf_ga, buf_z, K, C, S, s, 1e-3 / (1 - 0.5 ** step),
                               1 / math.sqrt(1 - 0.999 ** step), None, 0.5, 0.999, 1e-8, b_per_seg=B,
                               ga_seg_stride=per, z_seg_stride=per, n_seg=n_seg, g_out=g_out)
        assert rel(g_out, g_ref) < 1e-4
        # the moments are linear / quadratic in g: close to torch's; the parameter itself is sign-like in g at the first
        # steps (a ~0 gradient element may move by 2*lr either way), so it is checked against torch.optim.Adam fed
        # with the gradient this kernel formed
        st = opt.state[ref]
        assert rel(m, st['exp_avg']) < 1e-4 and rel(v, st['exp_avg_sq']) < 2e-4
        own.grad = g_out.clone()
        opt_own.step()
        assert torch.allclose(p, own.data, rtol=1e-5, atol=1e-6)
        assert torch.equal(img, o.prep_linear_weight(p, C, S))


def test_losses():
    o = ops()
    B = 16
    sr, sf = torch.randn(B, device='cuda'), torch.randn(B, device='cuda')
    out3, gr, gf = o.wloss(sr, sf, 1e-3)
    ref = -sr.mean() + sf.mean() + 1e-3 * (sr ** 2).mean()
    assert torch.allclose(out3, torch.stack([ref, sr.mean(), sf.mean()]), atol=1e-6)
    assert torch.allclose(gr, (-1 + 2e-3 * sr) / B, atol=1e-7) and torch.allclose(gf, torch.full_like(gf, 1 / B))
    out1, gf = o.gloss(sf)
    assert torch.allclose(out1[0], -sf.mean(), atol=1e-6) and torch.allclose(gf, torch.full_like(gf, -1 / B))
    g = (torch.randn(B, 32, 32, device='cuda') * 0.01).requires_grad_()
    pen_ref = 10 * (((0.5 * g.flatten(1).norm(dim=1)) - 1) ** 2).mean()
    gref, = torch.autograd.grad(pen_ref, g)
    pen, coeff = o.gp_loss(g.detach(), 0.5, 10.0)
    assert torch.allclose(pen[0], pen_ref, rtol=1e-5)
    # d pen / d g = coeff_b * norm_scale^2 * g
    assert torch.allclose(o.scale_rows(g.detach(), coeff, 0.25), gref, rtol=1e-4, atol=1e-8)
    pen2, coeff2 = o.gp_loss(g.detach(), 0.5, 10.0)
    assert torch.equal(pen, pen2) and torch.equal(coeff, coeff2)              # deterministic reduction
    # a sample with an exactly zero input gradient: torch's norm backward gives the zero subgradient there
    # (loss_functions.py:176); the coefficient must be 0, not -inf (which scale_rows would turn into NaN)
    gz = g.detach().clone()
    gz[3] = 0
    pen_z, coeff_z = o.gp_loss(gz, 0.5, 10.0)
    gzr = gz.clone().requires_grad_()
    pen_zr = 10 * (((0.5 * gzr.flatten(1).norm(dim=1)) - 1) ** 2).mean()
    gz_ref, = torch.autograd.grad(pen_zr, gzr)
    assert torch.allclose(pen_z[0], pen_zr, rtol=1e-5) and coeff_z[3].item() == 0.0
    out = o.scale_rows(gz, coeff_z, 0.25)
    assert torch.isfinite(out).all() and torch.allclose(out, gz_ref, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize('B,shape,L', [(16, (1, 64, 64), 512), (5, (1, 24, 40), 512), (64, (1, 16, 16), 512), (2, (1, 512, 512), 512)])
def test_similarity_loss(B, shape, L):
    """similarity_loss (reference loss_functions.py:185-205, verbatim formula below) as two kernels: per-chunk Gram
    matrices of the images, ordered reduction + cosine matrices + squared difference."""
    from neuron_gan_b200.loss_functions import similarity_loss
    g = torch.Generator(device='cuda').manual_seed(B)
    images = torch.rand(B, *shape, device='cuda', generator=g) * 2 - 1
    Z = torch.randn(B, L, device='cuda', generator=g)
    lam = 0.7
    im = images.view(B, -1).double()
    zz = Z.view(B, -1).double()
    im = im / im.norm(2, dim=1, keepdim=True)
    zz = zz / zz.norm(2, dim=1, keepdim=True)
    ref = lam * torch.pow(zz @ zz.t() - im @ im.t(), 2).sum() / (B * (B - 1))
    got = similarity_loss(images, Z, lam)
    assert got.dim() == 0 and abs(got.item() - ref.item()) <= 1e-4 * abs(ref.item()) + 1e-7, (got.item(), ref.item())
    assert similarity_loss(images, Z, lam).item() == got.item()                 # deterministic
    # inputs that require grad take the differentiable route (same value)
    zg = Z.clone().requires_grad_()
    d = similarity_loss(images, zg, lam)
    d.backward()
    assert abs(d.item() - ref.item()) <= 1e-4 * abs(ref.item()) + 1e-7 and zg.grad is not None


def test_adam_multi():
    o = ops()
    torch.manual_seed(0)
    shapes = [(128, 128, 3, 3), (16,), (1, 16, 1, 1), (1001,)]
    ps = [torch.randn(s, device='cuda') for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-3, betas=(0.5, 0.999))
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    shadow = torch.empty(ps[0].shape, dtype=torch.bfloat16, device='cuda')
    for step in range(1, 4):
        gs = [torch.randn_like(p) for p in ps]
        for q, g in zip(ref, gs):
            q.grad = g
        opt.step()
        ent = [dict(p=p, g=g, m=m, v=v, shadow=shadow if i == 0 else None, step_size=1e-3 / (1 - 0.5 ** step),
                    inv_bc2_sqrt=1 / math.sqrt(1 - 0.999 ** step)) for i, (p, g, m, v) in enumerate(zip(ps, gs, ms, vs))]
        o.adam_multi(ent, 0.5, 0.999, 1e-8)
    for p, q in zip(ps, ref):
        assert torch.allclose(p, q.data, rtol=1e-5, atol=1e-6)
    assert torch.equal(shadow, ps[0].to(torch.bfloat16))


def test_adam_refreshes_the_linear_operand_image():
    o = ops()
    C, S, K = 64, 4, 512
    p = torch.randn(C * S * S, K, device='cuda')
    g, m, v = torch.randn_like(p), torch.zeros_like(p), torch.zeros_like(p)
    img = torch.zeros(p.numel(), dtype=torch.bfloat16, device='cuda')
    o.adam_multi([dict(p=p, g=g, m=m, v=v, shadow=img, shadow_dims=(K, C, S * S), step_size=2e-4,
                       inv_bc2_sqrt=31.6)], 0.5, 0.999, 1e-8)
    assert torch.equal(img, o.prep_linear_weight(p, C, S))


@pytest.mark.parametrize('cin,cout', [(16, 16), (32, 64), (128, 128), (64, 128)])
def test_adam_refreshes_the_conv_operand_images(cin, cout):
    o = ops()
    p = torch.randn(cout, cin, 3, 3, device='cuda')
    g, m, v = torch.randn_like(p), torch.zeros_like(p), torch.zeros_like(p)
    n = p.numel()
    img = torch.zeros(2 * n, dtype=torch.bfloat16, device='cuda')
    o.adam_multi([dict(p=p, g=g, m=m, v=v, shadow=img, shadow_kind=2,
                       shadow_dims=(cin, cout, int(o.conv_weight_is_folded(cin, cout))), step_size=2e-4,
                       inv_bc2_sqrt=31.6)], 0.5, 0.999, 1e-8)
    fwd, dgrad = o.prep_conv_weight(p)
    assert torch.equal(img[:n], fwd) and torch.equal(img[n:], dgrad)


@pytest.mark.parametrize('B,cin,cout,H,W', [(2, 16, 16, 64, 64), (3, 32, 16, 40, 72), (2, 64, 64, 32, 32)])
@pytest.mark.parametrize('want_y', [True, False])
def test_conv3x3_fwd_with_fused_toimage(B, cin, cout, H, W, want_y):
    """conv + LeakyReLU + PixelNorm + ToImage in one kernel == torch fp32 reference of the same chain
    (models.py:203-204, 263-268, 141-149); y / r are optional outputs."""
    o = ops()
    x = rnd(B, cin, H, W, seed=31)
    w = rnd(cout, cin, 3, 3, seed=32, scale=GAIN / math.sqrt(cin * 9))
    tw = torch.randn(cout, device='cuda') * 0.3
    s = GAIN / math.sqrt(cin * 9)
    y_ref, r_ref = pn_ref(F.leaky_relu(F.conv2d(s * x, w, None, padding=1), LEAK))
    img_ref = torch.tanh((y_ref * tw.view(1, -1, 1, 1)).sum(1))
    w_fwd, _ = o.prep_conv_weight(w)
    y, r, img = o.conv3x3_fwd_toim(o.nchw_to_c8(x), w_fwd, None, s, LEAK, cout, tw, want_y=want_y, want_r=want_y)
    assert torch.allclose(img, img_ref, atol=1e-2) and rel(img, img_ref) < 1e-2
    if want_y:
        assert rel(o.c8_to_nchw(y), y_ref) < 1e-2 and rel(r, r_ref[:, 0]) < 1e-2
    else:
        assert y is None and r is None


def test_fused_toimage_is_refused_for_wide_layers():
    from neuron_gan_b200._lib import NganError
    o = ops()
    x = rnd(1, 128, 16, 16, seed=33)
    w = rnd(128, 128, 3, 3, seed=34)
    w_fwd, _ = o.prep_conv_weight(w)
    with pytest.raises(NganError):
        o.conv3x3_fwd_toim(o.nchw_to_c8(x), w_fwd, None, 1.0, LEAK, 128, torch.zeros(128, device='cuda'))


class _GuardedAlloc:
    """Stand-in for torch.empty / torch.empty_like while an op runs: every buffer the wrapper allocates (outputs AND
    scratch workspaces) sits between two 256-byte guard zones filled with a sentinel; `check()` asserts that no kernel
    wrote into them.  (compute-sanitizer is not available on the GPU pool; this is the out-of-bounds-write check for
    the kernels whose indexing was rewritten in round 2.)"""
    GUARD_BYTES = 256

    def __init__(self):
        self._empty, self._empty_like = torch.empty, torch.empty_like
        self.bufs = []

    def empty(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dtype = dtype or torch.float32
        n = 1
        for d in shape:
            n *= int(d)
        g = self.GUARD_BYTES // self._empty((), dtype=dtype).element_size()
        buf = self._empty(n + 2 * g, dtype=dtype, device=device)
        buf.fill_(1232.0)
        self.bufs.append((buf, g, n))
        return buf[g:g + n].view(shape)

    def empty_like(self, t, **kw):
        return self.empty(tuple(t.shape), dtype=kw.get('dtype', t.dtype), device=kw.get('device', t.device))

    def __enter__(self):
        torch.empty, torch.empty_like = self.empty, self.empty_like
        return self

    def __exit__(self, *exc):
        torch.empty, torch.empty_like = self._empty, self._empty_like

    def check(self, what):
        torch.cuda.synchronize()
        assert self.bufs, what
        for buf, g, n in self.bufs:
            lo, hi = buf[:g].float(), buf[g + n:].float()
            assert bool((lo == 1232.0).all()) and bool((hi == 1232.0).all()), f'{what}: write outside a {n}-element buffer'
        self.bufs = []


@pytest.mark.parametrize('B,C,R', [(2, 16, 128), (3, 32, 64), (2, 128, 16), (1, 16, 256)])
def test_no_out_of_bounds_writes(B, C, R):
    """Row-blocked resampling kernels, chunked 1x1-conv backward kernels and the float4 image kernels at shapes that
    take their vector / multi-row / multi-chunk paths: nothing is written outside the output and workspace buffers."""
    o = ops()
    torch.manual_seed(0)
    lo = o.nchw_to_c8(torch.randn(B, C, R // 2, R // 2, device='cuda'))
    hi = o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda'))
    r_lo = torch.rand(B, R // 2, R // 2, device='cuda') + 0.5
    r_hi = torch.rand(B, R, R, device='cuda') + 0.5
    img = torch.rand(B, R, R, device='cuda')
    img_lo = torch.rand(B, R // 2, R // 2, device='cuda')
    w = torch.randn(C, device='cuda')
    eps = torch.rand(B, device='cuda')
    gw, gb = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    with _GuardedAlloc() as ga:
        o.upsample2x(lo)
        ga.check('upsample2x')
        o.avgpool2(hi)
        ga.check('avgpool2')
        o.up2_bwd_pn_bwd(hi, lo, r_lo)
        ga.check('up2_bwd_pn_bwd')
        o.up2_bwd_pn_bwd(hi, lo, r_lo, extra_pre=img_lo, extra_w=w)
        ga.check('up2_bwd_pn_bwd + extra')
        gi = torch.empty(B, R, R, device='cuda')
        o.fromim_bwd(hi, img, w, gw, gb, g_img=gi, accumulate=False)
        ga.check('fromim_bwd')
        o.fromim_bwd(lo, img, w, gw, gb, gscale=0.25, unpool=True, g_img=gi, accumulate=True)
        ga.check('fromim_bwd unpool')
        o.fromim_dbl(img, hi, w, gw)
        ga.check('fromim_dbl')
        o.toim_bwd(img, img, hi, r_hi, w, gw, want_ga=True, want_gpre=True)
        ga.check('toim_bwd')
        o.pool_image(img)
        ga.check('pool_image')
        o.unpool_image(img_lo, 0.25)
        ga.check('unpool_image')
        o.up2_image(img_lo)
        ga.check('up2_image')
        o.up2_image_bwd(img, 0.5)
        ga.check('up2_image_bwd')
        o.interp_images(img, img.flip(0).contiguous(), eps)
        ga.check('interp_images')
        o.scale_rows(img, eps, 2.0)
        ga.check('scale_rows')
        o.lerp(img, img.flip(0).contiguous(), 0.3)
        ga.check('lerp')
        o.axpby(img, 0.5, img, 0.25)
        ga.check('axpby')
