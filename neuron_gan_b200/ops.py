"""Tensor-level wrappers over the C ABI (include/ngan_b200.h).  torch is used for device memory and streams
only: every function validates its tensors, takes raw device pointers and launches on the current stream.

Feature maps are "c8" tensors: torch.bfloat16, shape [B, C//8, H, W, 8], contiguous (see csrc/common.cuh).
"""
import ctypes

import torch

from . import _lib

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.NganError('neuron_gan_b200 kernels need CUDA tensors (there is no CPU fallback)')
    if not t.is_contiguous():
        raise _lib.NganError('non-contiguous tensor passed to a neuron_gan_b200 kernel')
    if dtype is not None and t.dtype != dtype:
        raise _lib.NganError(f'expected {dtype}, got {t.dtype}')
    return ctypes.c_void_p(t.data_ptr())


def _scalar(v):
    """A scalar kernel argument: a float, or (float, 1-element fp32 device tensor) meaning their product with the
    tensor read by the kernel (the `dyn` pointers of include/ngan_b200.h) -- used for the fade-in coefficient alpha
    so that a captured graph follows it."""
    if isinstance(v, tuple):
        c, t = v
        assert t.numel() == 1
        return float(c), _p(t, F32)
    return float(v), None


def scalar_mul(c, v):
    """float * scalar-argument (see _scalar)."""
    if isinstance(v, tuple):
        return (float(c) * v[0], v[1])
    return float(c) * v


def c8_empty(B, C, H, W, device):
    return torch.empty((B, C // 8, H, W, 8), dtype=BF16, device=device)


def c8_dims(t):
    B, nch, H, W, e = t.shape
    assert e == 8 and t.dtype == BF16
    return B, nch * 8, H, W


# ------------------------------------------------------------------------------------------ layout
def nchw_to_c8(x):
    B, C, H, W = x.shape
    out = c8_empty(B, C, H, W, x.device)
    _lib.call('ngan_nchw_to_c8', _p(x, F32), _p(out), B, C, H, W, _stream())
    return out


def c8_to_nchw(x):
    B, C, H, W = c8_dims(x)
    out = torch.empty((B, C, H, W), dtype=F32, device=x.device)
    _lib.call('ngan_c8_to_nchw', _p(x, BF16), _p(out), B, C, H, W, _stream())
    return out


# ------------------------------------------------------------------------------------------ conv 3x3
def conv_weight_is_folded(cin, cout):
    return bool(_lib.call('ngan_conv_weight_is_folded', cin, cout))


def prep_conv_weight(w, w_fwd=None, w_dgrad=None):
    """w: fp32 [cout, cin, 3, 3] -> (fwd image, dgrad image), bf16 flat buffers of cout*cin*9 elements (views of one
    2*cout*cin*9 buffer when allocated here: the layout the Adam kernel refreshes in place)."""
    cout, cin = w.shape[0], w.shape[1]
    if w_fwd is None and w_dgrad is None:
        both = torch.empty(2 * cout * cin * 9, dtype=BF16, device=w.device)
        w_fwd, w_dgrad = both[:cout * cin * 9], both[cout * cin * 9:]
    if w_fwd is None:
        w_fwd = torch.empty(cout * cin * 9, dtype=BF16, device=w.device)
    if w_dgrad is None:
        w_dgrad = torch.empty(cout * cin * 9, dtype=BF16, device=w.device)
    _lib.call('ngan_prep_conv_weight', _p(w, F32), _p(w_fwd, BF16), _p(w_dgrad, BF16), cin, cout, _stream())
    return w_fwd, w_dgrad


def conv3x3_fwd(x, w_fwd, bias, scale, leak, cout, want_r=True):
    B, cin, H, W = c8_dims(x)
    y = c8_empty(B, cout, H, W, x.device)
    r = torch.empty((B, H, W), dtype=F32, device=x.device) if want_r else None
    _lib.call('ngan_conv3x3_fwd', _p(x, BF16), _p(w_fwd, BF16), _p(bias, F32), scale, leak, _p(y), _p(r), B, cin,
              cout, H, W, _stream())
    return y, r


def conv3x3_fwd_toim(x, w_fwd, bias, scale, leak, cout, toim_w, want_y=True, want_r=True, img_out=None):
    """conv + LeakyReLU + PixelNorm with ToImage fused: returns (y or None, r or None, img [B, H, W] fp32)."""
    B, cin, H, W = c8_dims(x)
    y = c8_empty(B, cout, H, W, x.device) if want_y else None
    r = torch.empty((B, H, W), dtype=F32, device=x.device) if want_r else None
    img = torch.empty((B, H, W), dtype=F32, device=x.device) if img_out is None else img_out
    assert img.shape == (B, H, W) and img.dtype in (F32, BF16)       # bf16 images: generator-only inference
    _lib.call('ngan_conv3x3_fwd_toim', _p(x, BF16), _p(w_fwd, BF16), _p(bias, F32), scale, leak, _p(y), _p(r),
              _p(toim_w, F32), _p(img), int(img.dtype == BF16), B, cin, cout, H, W, _stream())
    return y, r, img


def f32_to_bf16(src, out):
    assert src.dtype == F32 and out.dtype == BF16 and src.numel() == out.numel()
    _lib.call('ngan_f32_to_bf16', _p(src, F32), _p(out, BF16), src.numel(), _stream())
    return out


def conv3x3_dgrad(ga, w_dgrad, scale, cin):
    B, cout, H, W = c8_dims(ga)
    gx = c8_empty(B, cin, H, W, ga.device)
    _lib.call('ngan_conv3x3_dgrad', _p(ga, BF16), _p(w_dgrad, BF16), scale, _p(gx), B, cin, cout, H, W, _stream())
    return gx


def conv3x3_dgrad_pn(ga, w_dgrad, scale, leak, y_prev, r_prev, addin=None, want_gy=False):
    B, cout, H, W = c8_dims(ga)
    cin = c8_dims(y_prev)[1]
    ga_prev = torch.empty_like(y_prev)
    gy = torch.empty_like(y_prev) if want_gy else None
    _lib.call('ngan_conv3x3_dgrad_pn', _p(ga, BF16), _p(w_dgrad, BF16), scale, leak, _p(y_prev, BF16),
              _p(r_prev, F32), _p(addin, BF16) if addin is not None else None, _p(ga_prev), _p(gy), B, cin, cout, H,
              W, _stream())
    return ga_prev, gy


def conv3x3_dbl(ghat_x, w_fwd, scale, leak, y, r, gy):
    B, cin, H, W = c8_dims(ghat_x)
    cout = c8_dims(y)[1]
    ghat_y = torch.empty_like(y)
    ahat = torch.empty_like(y)
    _lib.call('ngan_conv3x3_dbl', _p(ghat_x, BF16), _p(w_fwd, BF16), scale, leak, _p(y, BF16), _p(r, F32),
              _p(gy, BF16), _p(ghat_y), _p(ahat), B, cin, cout, H, W, _stream())
    return ghat_y, ahat


def _workspace(nbytes, device):
    """fp32 scratch from torch's caching allocator, allocated on the CURRENT stream (so a kernel launched on a side
    stream gets a block that only that stream's later allocations can reuse)."""
    return torch.empty((int(nbytes) + 3) // 4, dtype=F32, device=device)


def _pixel_ws(B, C, H, W, device):
    return _workspace(_lib.call('ngan_pixel_reduction_workspace_bytes', B, C, H, W), device)


def conv3x3_wgrad(x, ga, scale, dw, accumulate=True):
    """dw (fp32 [cout, cin, 3, 3]) (+)= scale * corr(x, ga); deterministic (per-CTA partials added in order)"""
    B, cin, H, W = c8_dims(x)
    cout = c8_dims(ga)[1]
    assert dw.numel() == cout * cin * 9
    ws = _workspace(_lib.call('ngan_conv3x3_wgrad_workspace_bytes', B, cin, cout, H, W), x.device)
    _lib.call('ngan_conv3x3_wgrad', _p(x, BF16), _p(ga, BF16), scale, _p(dw, F32), int(accumulate), _p(ws, F32), B,
              cin, cout, H, W, _stream())


def bias_grad(ga, gb, accumulate=True):
    B, C, H, W = c8_dims(ga)
    ws = _pixel_ws(B, C, H, W, ga.device)
    _lib.call('ngan_bias_grad', _p(ga, BF16), _p(gb, F32), int(accumulate), _p(ws, F32), B, C, H, W, _stream())


def reduce_partials(partials, n_partials, n, ld, out, scale=1.0, accumulate=False):
    _lib.call('ngan_reduce_partials', _p(partials, F32), n_partials, n, ld, scale, _p(out, F32), int(accumulate),
              _stream())


def sum_slots(slots, out, scale=1.0):
    """out = scale * (slots[0] + slots[1] + ...) in slot order; slots: [n_slots, n] fp32, rows `slots.stride(0)` apart"""
    n_slots, n = slots.shape
    if slots.stride(1) != 1 or not slots.is_cuda or slots.dtype != F32:
        raise _lib.NganError('sum_slots needs fp32 CUDA rows with unit inner stride')
    _lib.call('ngan_sum_slots', ctypes.c_void_p(slots.data_ptr()), n_slots, n, slots.stride(0), float(scale),
              _p(out, F32), _stream())
    return out


def memset_zero(t):
    """cudaMemsetAsync on the current stream (no ATen fill kernel inside the captured iteration)"""
    if not t.is_contiguous():
        raise _lib.NganError('memset_zero needs a contiguous tensor')
    _lib.call('ngan_memset', ctypes.c_void_p(t.data_ptr()), 0, t.numel() * t.element_size(), _stream())
    return t


# ------------------------------------------------------------------------------------------ resampling
def upsample2x(x):
    B, C, H, W = c8_dims(x)
    out = c8_empty(B, C, 2 * H, 2 * W, x.device)
    _lib.call('ngan_upsample2x', _p(x, BF16), _p(out), B, C, H, W, _stream())
    return out


def avgpool2(x):
    B, C, H, W = c8_dims(x)
    out = c8_empty(B, C, H // 2, W // 2, x.device)
    _lib.call('ngan_avgpool2', _p(x, BF16), _p(out), B, C, H, W, _stream())
    return out


def pn_bwd(g, y, r, gscale=1.0, unpool=False, addin=None, want_gy=False, leak=0.2):
    B, C, H, W = c8_dims(y)
    ga = torch.empty_like(y)
    gy = torch.empty_like(y) if want_gy else None
    gs, dyn = _scalar(gscale)
    _lib.call('ngan_pn_bwd', _p(g, BF16), int(unpool), gs, dyn, _p(y, BF16), _p(r, F32),
              _p(addin, BF16) if addin is not None else None, _p(ga), _p(gy), leak, B, C, H, W, _stream())
    return ga, gy


def up2_bwd_pn_bwd(g_up, y, r, extra_pre=None, extra_w=None, leak=0.2):
    B, C, H, W = c8_dims(y)
    ga = torch.empty_like(y)
    _lib.call('ngan_up2_bwd_pn_bwd', _p(g_up, BF16), _p(y, BF16), _p(r, F32), _p(extra_pre, F32), _p(extra_w, F32),
              _p(ga), leak, B, C, H, W, _stream())
    return ga


# ------------------------------------------------------------------------------------------ 1-channel images
def pool_image(x):
    B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
    out = torch.empty((B, H // 2, W // 2), dtype=F32, device=x.device)
    _lib.call('ngan_pool_image', _p(x, F32), _p(out), B, H, W, _stream())
    return out


def unpool_image(g, scale):
    B, h, w = g.shape
    out = torch.empty((B, 2 * h, 2 * w), dtype=F32, device=g.device)
    _lib.call('ngan_unpool_image', _p(g, F32), _p(out), scale, B, 2 * h, 2 * w, _stream())
    return out


def up2_image(x):
    B, H, W = x.shape
    out = torch.empty((B, 2 * H, 2 * W), dtype=F32, device=x.device)
    _lib.call('ngan_up2_image', _p(x, F32), _p(out), B, H, W, _stream())
    return out


def up2_image_bwd(g, scale=1.0):
    B, H2, W2 = g.shape
    out = torch.empty((B, H2 // 2, W2 // 2), dtype=F32, device=g.device)
    sc, dyn = _scalar(scale)
    _lib.call('ngan_up2_image_bwd', _p(g, F32), _p(out), sc, dyn, B, H2 // 2, W2 // 2, _stream())
    return out


def lerp(a, b, alpha, out=None):
    out = torch.empty_like(a) if out is None else out
    al, dyn = _scalar(alpha)
    _lib.call('ngan_lerp', _p(a, F32), _p(b, F32), al, dyn, _p(out), a.numel(), _stream())
    return out


def axpby(a, ca, b=None, cb=0.0, out=None):
    out = torch.empty_like(a) if out is None else out
    _lib.call('ngan_axpby', _p(a, F32), ca, _p(b, F32), cb, _p(out), a.numel(), _stream())
    return out


def interp_images(real, fake, eps):
    out = torch.empty_like(real)
    B = real.shape[0]
    _lib.call('ngan_interp_images', _p(real, F32), _p(fake, F32), _p(eps, F32), _p(out), B, real.numel() // B,
              _stream())
    return out


def scale_rows(x, coeff, scale=1.0):
    out = torch.empty_like(x)
    B = x.shape[0]
    _lib.call('ngan_scale_rows', _p(x, F32), _p(coeff, F32), scale, _p(out), B, x.numel() // B, _stream())
    return out


# ------------------------------------------------------------------------------------------ FromImage / ToImage
def fromim_fwd(xp, w, b):
    B, H, W = xp.shape
    C = w.numel()
    out = c8_empty(B, C, H, W, xp.device)
    _lib.call('ngan_fromim_fwd', _p(xp, F32), _p(w, F32), _p(b, F32), _p(out), B, C, H, W, _stream())
    return out


def d_fade_fwd(y_end, xp, w_old, b_old, alpha):
    B, C, H, W = c8_dims(y_end)
    out = torch.empty_like(y_end)
    al, dyn = _scalar(alpha)
    _lib.call('ngan_d_fade_fwd', _p(y_end, BF16), _p(xp, F32), _p(w_old, F32), _p(b_old, F32), al, dyn, _p(out), B, C,
              H, W, _stream())
    return out


def fromim_bwd(g, xp, w, gw, gb, gscale=1.0, unpool=False, g_img=None, accumulate=False, grad_accumulate=True):
    """accumulate: g_img += (else =); grad_accumulate: gw / gb += (else =)"""
    B, H, W = xp.shape
    C = w.numel()
    gs, dyn = _scalar(gscale)
    want = gw is not None or gb is not None
    ws = _pixel_ws(B, C, H, W, xp.device) if want else None
    _lib.call('ngan_fromim_bwd', _p(g, BF16), int(unpool), gs, dyn, _p(xp, F32), _p(w, F32), _p(gw, F32), _p(gb, F32),
              int(grad_accumulate), _p(ws, F32), _p(g_img, F32), int(accumulate), B, C, H, W, _stream(),
              launches=1 + (1 if (gw is not None and gb is not None) else (gw is not None) + (gb is not None)))


def fromim_dbl(ghat_xp, g, w, what, in_scale=1.0, gscale=1.0, unpool=False, want_out=True, grad_accumulate=True):
    B, H, W = ghat_xp.shape
    C = w.numel()
    out = c8_empty(B, C, H, W, ghat_xp.device) if want_out else None
    gs, dyn = _scalar(gscale)
    ws = _pixel_ws(B, C, H, W, ghat_xp.device) if what is not None else None
    _lib.call('ngan_fromim_dbl', _p(ghat_xp, F32), in_scale, _p(g, BF16), int(unpool), gs, dyn, _p(w, F32), _p(out),
              _p(what, F32), int(grad_accumulate), _p(ws, F32), B, C, H, W, _stream(),
              launches=1 + (what is not None))
    return out


def toim_fwd(y, w, out=None):
    B, C, H, W = c8_dims(y)
    img = torch.empty((B, H, W), dtype=F32, device=y.device) if out is None else out
    assert img.shape == (B, H, W) and img.dtype == F32
    _lib.call('ngan_toim_fwd', _p(y, BF16), _p(w, F32), _p(img), B, C, H, W, _stream())
    return img


def toim_bwd(g_img, img, y, r, w, gw, gscale=1.0, want_ga=True, want_gpre=False, leak=0.2, grad_accumulate=True):
    B, C, H, W = c8_dims(y)
    ga = torch.empty_like(y) if want_ga else None
    gpre = torch.empty((B, H, W), dtype=F32, device=y.device) if want_gpre else None
    gs, dyn = _scalar(gscale)
    ws = _pixel_ws(B, C, H, W, y.device) if gw is not None else None
    _lib.call('ngan_toim_bwd', _p(g_img, F32), gs, dyn, _p(img, F32), _p(y, BF16), _p(r, F32), _p(w, F32), _p(ga),
              _p(gpre), _p(gw, F32), int(grad_accumulate), _p(ws, F32), leak, B, C, H, W, _stream(),
              launches=1 + (gw is not None))
    return ga, gpre


# ------------------------------------------------------------------------------------------ critic head
def head_fwd(y, w, bias, scale):
    B, C, H, W = c8_dims(y)
    score = torch.empty((B,), dtype=F32, device=y.device)
    _lib.call('ngan_head_fwd', _p(y, BF16), _p(w, F32), _p(bias, F32), scale, _p(score), B, C, H, _stream())
    return score


def head_bwd_pn(gout, w, scale, y, r, want_gy=False, leak=0.2):
    B, C, H, W = c8_dims(y)
    ga = torch.empty_like(y)
    gy = torch.empty_like(y) if want_gy else None
    _lib.call('ngan_head_bwd_pn', _p(gout, F32), _p(w, F32), scale, _p(y, BF16), _p(r, F32), _p(ga), _p(gy), leak, B,
              C, H, _stream())
    return ga, gy


def head_wgrad(t, coeff, scale, gw, gb=None, accumulate=True):
    """gw (+)= scale * sum_b coeff[b] * t[b]; gb (the head's bias gradient, optional) (+)= sum_b coeff[b]."""
    B, C, H, W = c8_dims(t)
    _lib.call('ngan_head_wgrad', _p(t, BF16), _p(coeff, F32), scale, _p(gw, F32), _p(gb, F32), int(accumulate), B, C,
              H, _stream())


# ------------------------------------------------------------------------------------------ generator stem
def prep_linear_weight(w, C, S, out=None):
    """w: fp32 [C*S*S, K] -> bf16 operand image, flat [S*S][K/8][C][8] (see csrc/linear.cu)."""
    K = w.shape[1]
    assert w.shape[0] == C * S * S
    out = torch.empty(w.numel(), dtype=BF16, device=w.device) if out is None else out
    _lib.call('ngan_prep_linear_weight', _p(w, F32), _p(out, BF16), K, C, S, _stream())
    return out


def linear_image_to_matrix(img, K, C, S):
    """Inverse permutation of prep_linear_weight (tests / debugging): image -> [C*S*S, K] bf16."""
    return img.view(S * S, K // 8, C, 8).permute(2, 0, 1, 3).reshape(C * S * S, K)


def linear_fwd(z, w_img, scale, leak, C, S, want_r=True):
    B, K = z.shape
    y = c8_empty(B, C, S, S, z.device)
    r = torch.empty((B, S, S), dtype=F32, device=z.device) if want_r else None
    ws = torch.empty(_lib.call('ngan_linear_fwd_workspace_bytes', B, K), dtype=torch.uint8, device=z.device)
    _lib.call('ngan_linear_fwd', _p(z, F32), _p(w_img, BF16), scale, leak, _p(y), _p(r), _p(ws), B, K, C, S, _stream())
    return y, r


def linear_wgrad(ga, z, scale, dw, accumulate=True):
    B, C, S, _ = c8_dims(ga)
    _lib.call('ngan_linear_wgrad', _p(ga, BF16), _p(z, F32), scale, _p(dw, F32), int(accumulate), B, z.shape[1], C, S,
              _stream())


# ------------------------------------------------------------------------------------------ losses
def wloss(s_real, s_fake, drift, gscale=1.0, want_grads=True):
    B = s_real.numel()
    out3 = torch.empty(3, dtype=F32, device=s_real.device)
    g_real = torch.empty(B, dtype=F32, device=s_real.device) if want_grads else None
    g_fake = torch.empty(B, dtype=F32, device=s_real.device) if want_grads else None
    _lib.call('ngan_wloss', _p(s_real, F32), _p(s_fake, F32), drift, _p(out3), _p(g_real), _p(g_fake), gscale, B,
              _stream())
    return out3, g_real, g_fake


def wloss_into(s_real, s_fake, drift, g_real, g_fake, gscale=1.0):
    """wloss writing the per-sample score gradients into caller-provided (contiguous) slices."""
    out3 = torch.empty(3, dtype=F32, device=s_real.device)
    _lib.call('ngan_wloss', _p(s_real, F32), _p(s_fake, F32), drift, _p(out3), _p(g_real, F32), _p(g_fake, F32),
              gscale, s_real.numel(), _stream())
    return out3, g_real, g_fake


def gloss(s_fake, gscale=1.0, want_grads=True):
    B = s_fake.numel()
    out1 = torch.empty(1, dtype=F32, device=s_fake.device)
    g_fake = torch.empty(B, dtype=F32, device=s_fake.device) if want_grads else None
    _lib.call('ngan_gloss', _p(s_fake, F32), _p(out1), _p(g_fake), gscale, B, _stream())
    return out1, g_fake


def gp_loss(g, norm_scale, lam, gscale=1.0):
    B = g.shape[0]
    pen = torch.empty(1, dtype=F32, device=g.device)
    coeff = torch.empty(B, dtype=F32, device=g.device)
    ws = _workspace(_lib.call('ngan_gp_loss_workspace_bytes', B), g.device)
    _lib.call('ngan_gp_loss', _p(g, F32), norm_scale, lam, _p(pen), _p(coeff), gscale, _p(ws, F32), B, g.numel() // B,
              _stream())
    return pen, coeff


def similarity_loss(images, z, lam):
    """images [B, ...] fp32, z [B, L] fp32 -> [1] fp32 (reference loss_functions.py:185-205)"""
    B = images.shape[0]
    per = images.numel() // B
    out = torch.empty(1, dtype=F32, device=images.device)
    ws = _workspace(_lib.call('ngan_similarity_loss_workspace_bytes', B, per), images.device)
    _lib.call('ngan_similarity_loss', _p(images, F32), _p(z, F32), float(lam), _p(ws, F32), _p(out), B, per,
              z.numel() // B, _stream())
    return out


def pack_stats(out3, out1, pen, stats):
    _lib.call('ngan_pack_stats', _p(out3, F32), _p(out1, F32), _p(pen, F32), _p(stats, F32), _stream())
    return stats


# ------------------------------------------------------------------------------------------ Adam
def adam_multi(entries, beta1, beta2, eps):
    """entries: list of dicts(p, g, m, v, shadow|None, step_size, inv_bc2_sqrt, dyn|None)"""
    n = len(entries)
    if n == 0:
        return
    arr = (_lib.AdamTensor * n)()
    for i, e in enumerate(entries):
        arr[i].p = e['p'].data_ptr()
        arr[i].g = e['g'].data_ptr()
        arr[i].m = e['m'].data_ptr()
        arr[i].v = e['v'].data_ptr()
        arr[i].shadow_bf16 = e['shadow'].data_ptr() if e.get('shadow') is not None else None
        arr[i].n = e['p'].numel()
        arr[i].step_size = e['step_size']
        arr[i].inv_bc2_sqrt = e['inv_bc2_sqrt']
        arr[i].dyn = e['dyn'].data_ptr() if e.get('dyn') is not None else None
        arr[i].shadow_k, arr[i].shadow_c, arr[i].shadow_ss = e.get('shadow_dims') or (0, 0, 0)
        arr[i].shadow_kind = e.get('shadow_kind', 1 if e.get('shadow_dims') else 0)
    _lib.call('ngan_adam_multi', ctypes.cast(arr, ctypes.c_void_p), n, beta1, beta2, eps, _stream())


def adam_linear_factored(p, m, v, shadow, ga, z, K, C, S, gscale, step_size, inv_bc2_sqrt, dyn, beta1, beta2, eps,
                         b_per_seg=None, ga_seg_stride=0, z_seg_stride=0, n_seg=1, g_out=None):
    """Adam on the Linear weight with the gradient formed from its factors (see ngan_adam_linear_factored).
    ga: c8 gradient at the stem's pre-activation, z: latents [B, K]; with n_seg > 1 both are all-gathered buffers of
    n_seg per-rank segments of b_per_seg samples, *_seg_stride bytes apart."""
    b_per_seg = b_per_seg if b_per_seg is not None else z.shape[0]
    _lib.call('ngan_adam_linear_factored', _p(p, F32), _p(m, F32), _p(v, F32), _p(shadow, BF16), _p(g_out, F32),
              ctypes.c_void_p(ga.data_ptr()), ctypes.c_void_p(z.data_ptr()), b_per_seg * n_seg, b_per_seg,
              int(ga_seg_stride), int(z_seg_stride), K, C, S, float(gscale), float(step_size), float(inv_bc2_sqrt),
              _p(dyn, F32), beta1, beta2, eps, _stream())


def linear_wgrad_factored(ga, z, K, C, S, gscale, dw, b_per_seg, n_seg=1, ga_seg_stride=0, z_seg_stride=0):
    """dW = gscale * ga^T z over all samples of an (all-gathered) global batch, on tensor cores (z = hi + lo in bf16):
    the gradient-only mode of ngan_adam_linear_factored.  Overwrites dw."""
    _lib.call('ngan_adam_linear_factored', None, None, None, None, _p(dw, F32), ctypes.c_void_p(ga.data_ptr()),
              ctypes.c_void_p(z.data_ptr()), b_per_seg * n_seg, b_per_seg, int(ga_seg_stride), int(z_seg_stride), K, C,
              S, float(gscale), 0.0, 0.0, None, 0.0, 0.0, 0.0, _stream())


# ------------------------------------------------------------------------------------------ image pipeline
def augment_batch(canvases, src_index, params, tap_first, tap_count, tap_weight, out, crop, workspace=None):
    """canvases [N, P, P] f32, src_index [b] i32, params [b, 16] f32, taps for crop -> R, out [b, 1, R, R] f32
    (data/NeuronDataset.py:170-205 for one whole batch; see ngan_augment_batch in include/ngan_b200.h)."""
    b, R = out.shape[0], out.shape[-1]
    N, P, P2 = canvases.shape
    assert P == P2 and out.shape == (b, 1, R, R) and params.shape == (b, 16) and src_index.shape == (b,)
    assert tap_first.shape == (R,) and tap_count.shape == (R,) and tap_weight.shape[0] == R
    need = _lib.call('ngan_augment_workspace_bytes', b, P, crop)
    if workspace is None or workspace.numel() * 4 < need:
        workspace = torch.empty((need + 3) // 4, dtype=F32, device=out.device)
    _lib.call('ngan_augment_batch', _p(canvases, F32), _p(src_index, torch.int32), _p(params, F32),
              _p(tap_first, torch.int32), _p(tap_count, torch.int32), _p(tap_weight, F32), tap_weight.shape[1],
              _p(workspace, F32), _p(out, F32), b, P, crop, R, _stream())
    return out
