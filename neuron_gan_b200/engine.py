"""Network-level orchestration of the sm_100a kernels: forward, backward and the gradient-penalty double
backward of Generator_PG / Discriminator_PG, written out by hand (no autograd inside).

The reference gets all of this from ATen + autograd (reference models.py:344-353, 516-524 for the forwards;
loss_functions.py:175 `autograd.grad(..., create_graph=True)` + train.py:365 `.backward()` for the rest).
Here each pass is an explicit kernel sequence over saved activations:

  * a conv + LeakyReLU + PixelNorm stage saves only its output y (bf16, it is the next stage's input anyway)
    and the per-pixel PixelNorm scale r; the LeakyReLU mask is recovered from sign(y);
  * backward: the data-gradient conv of stage l fuses the PixelNorm/LeakyReLU backward of stage l-1 in its
    epilogue whenever no resampling sits between them, otherwise a pointwise kernel does the resample adjoint
    + PixelNorm backward in one pass;
  * double backward (D only): sweep 1 runs the "conv of the cotangent" kernels from the image side up to the
    head and leaves, per stage, the cotangent to inject at its pre-activation; sweep 2 is the ordinary
    backward with those injections.  Formulas: SURVEY.md section 8a row 3.

Parameter gradients go to the fp32 tensors of a `sink` (Sink: id(param) -> tensor).  Every pass (a backward, a sweep of
the double backward) touches each parameter with exactly ONE kernel, and no kernel uses atomics, so a sink with
accumulate=False needs no zeroing and its content is bit-reproducible.  Passes that may run concurrently (the
Wasserstein backward and the two sweeps of the penalty's double backward all contribute to every critic weight)
must get SEPARATE sinks, which the caller adds afterwards (TrainStep: three slots + ops.sum_slots).
"""
import os
import weakref
from types import SimpleNamespace

import torch

from . import ops

F32 = torch.float32

# ---------------------------------------------------------------------------------------------------------
# bf16 operand images of the fp32 master weights, refreshed when the master changes
# ---------------------------------------------------------------------------------------------------------
_weight_cache = {}          # id(param) -> entry; entries die with their parameter


def _cache_get(w):
    ent = _weight_cache.get(id(w))
    return ent if ent is not None and ent['ref']() is w else None


def _cache_set(w, ent):
    if _cache_get(w) is None:
        weakref.finalize(w, _weight_cache.pop, id(w), None)
    ent['ref'] = weakref.ref(w)
    _weight_cache[id(w)] = ent


def _wkey(w):
    return (w._version, getattr(w, '_ngan_epoch', 0), getattr(w, '_ngan_ext', 0), w.data_ptr())


def conv_images(conv):
    """(forward image, data-gradient image) of a Conv2d_normalized 3x3 weight."""
    w = conv.weight
    ent = _cache_get(w)
    key = _wkey(w)
    if ent is None or ent['key'] != key:
        bufs = (ent['fwd'], ent['dgrad']) if ent is not None and ent['fwd'].device == w.device else (None, None)
        fwd, dgrad = ops.prep_conv_weight(w.detach(), *bufs)
        cout, cin = w.shape[0], w.shape[1]
        # fwd and dgrad are the two halves of one buffer: FusedAdam rewrites both while it updates the master
        ent = {'key': key, 'fwd': fwd, 'dgrad': dgrad, 'shadow': fwd, 'shadow_kind': 2,
               'shadow_dims': (cin, cout, int(ops.conv_weight_is_folded(cin, cout)))}
        _cache_set(w, ent)
    return ent['fwd'], ent['dgrad']


def prepare_weights(net):
    """Refresh (on the current stream) the bf16 operand images of every 3x3 conv of `net` whose master weight
    changed.  Called before work on `net` is forked to several streams, so that no stream depends on a
    preparation kernel another stream launched."""
    for m in net.modules():
        if isinstance(m, torch.nn.Conv2d) and hasattr(m, 'scale_value') and tuple(m.kernel_size) == (3, 3):
            conv_images(m)


def _linear_dims(lin):
    """(K, C, S) of the generator's Linear_normalized: latent -> C*S*S features (models.py:299-302)."""
    return lin._ngan_dims


def linear_shadow(lin, allocate_only=False):
    """bf16 operand image of the generator's Linear_normalized weight, [S*S][K/8][C][8] (csrc/linear.cu)."""
    w = lin.weight
    ent = _cache_get(w)
    key = _wkey(w)
    K, C, S = _linear_dims(lin)
    if ent is None or ent['shadow'].device != w.device:
        ent = {'key': None, 'shadow': torch.empty(w.numel(), dtype=torch.bfloat16, device=w.device),
               'shadow_dims': (K, C, S * S), 'shadow_kind': 1}
        _cache_set(w, ent)
    if allocate_only:
        return ent
    if ent['key'] != key:
        ops.prep_linear_weight(w.detach(), C, S, ent['shadow'])
        ent['key'] = key
    return ent['shadow']


def invalidate(net_or_params):
    """Declare the fp32 master weights of a network (or an iterable of parameters) changed from outside the fused
    optimiser.  The bf16 operand images (conv_images / linear_shadow) and TrainStep's decision to replay a captured
    graph key on `p._version`, which in-place writes through `p.data` (EMA weight swaps, weight clipping,
    re-initialisation after a first forward: `p.data.copy_()`, `p.data.clamp_()` ...) do NOT bump -- such writes must
    be followed by engine.invalidate(net), otherwise the kernels keep computing with the old images.  Writes through
    the parameter itself under no_grad (`p.copy_()`, load_state_dict, optimizer.step) need nothing."""
    params = net_or_params.parameters() if hasattr(net_or_params, 'parameters') else net_or_params
    for p in params:
        p._ngan_ext = getattr(p, '_ngan_ext', 0) + 1      # (not _ngan_epoch: that one counts the fused optimiser's steps)


def mark_updated(param, shadow_is_fresh=False):
    """Called by the fused optimiser after it changed `param` through a raw pointer."""
    param._ngan_epoch = getattr(param, '_ngan_epoch', 0) + 1
    if shadow_is_fresh:
        ent = _cache_get(param)
        if ent is not None and 'shadow' in ent:
            ent['key'] = _wkey(param)


# ---------------------------------------------------------------------------------------------------------
# Side stream for the weight-gradient kernels.  A wgrad depends only on a saved activation and on the gradient
# the data-gradient chain has just produced, and nothing downstream waits for it until the optimiser, so it
# runs beside the chain instead of in it (the narrow layers' wgrads use a fraction of the SMs).  Works eagerly
# (events) and under CUDA-graph capture (fork/join edges).
# ---------------------------------------------------------------------------------------------------------
class _Side:
    streams = []       # torch.cuda.Stream pool, created on first use
    n_streams = 6      # the low-resolution wgrads (kernel + dependent reduction) are latency-bound on a few dozen CTAs:
                       # several run side by side (measured, graph replay: 16x16 phase 0.667 -> 0.625 ms from 3 to 6
                       # streams, 64x64 2.08 -> 2.06, 512x512 unchanged)
    enabled = True
    keep = []          # tensors a pending side-stream kernel reads: kept alive until side_join()
    used = set()       # indices of pool streams with work since the last join
    rr = 0


def _on_side(fn, *tensors):
    """Run fn() -- a parameter-gradient kernel whose only output is its own slice of a gradient buffer, which nothing
    reads before side_join() -- on a side stream, after everything already queued on the current stream; `tensors`
    are its inputs (kept alive)."""
    if not _Side.enabled:
        fn()
        return
    if not _Side.streams:
        _Side.streams = [torch.cuda.Stream() for _ in range(_Side.n_streams)]
    i = _Side.rr % len(_Side.streams)
    _Side.rr += 1
    side = _Side.streams[i]
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    _Side.keep.append(tensors)
    _Side.used.add(i)


class Sink(dict):
    """id(param) -> fp32 gradient tensor.  accumulate=True: kernels add to the tensors (the caller zeroed them and no
    other pass writes them concurrently); accumulate=False: kernels overwrite them."""

    def __init__(self, *args, accumulate=True, **kw):
        super().__init__(*args, **kw)
        self.accumulate = accumulate


def _acc(sink):
    return getattr(sink, 'accumulate', True)


def _wgrad(x, ga, scale, dw, acc=True):
    """conv3x3_wgrad on a side stream (nothing downstream waits for it until the optimiser)."""
    if dw is not None:
        _on_side(lambda: ops.conv3x3_wgrad(x, ga, scale, dw, accumulate=acc), x, ga)


def side_join():
    """Make the current stream wait for every side-stream wgrad issued so far (call before reading gradients)."""
    if _Side.used:
        cur = torch.cuda.current_stream()
        for i in sorted(_Side.used):
            cur.wait_stream(_Side.streams[i])
        _Side.keep.clear()
        _Side.used.clear()


def _sink_get(sink, p):
    return None if sink is None else sink[id(p)]


def _flat(w):
    return w.detach().reshape(-1)


# ---------------------------------------------------------------------------------------------------------
# Generator
# ---------------------------------------------------------------------------------------------------------
FUSE_TOIMAGE = os.environ.get('NGAN_NO_FUSED_TOIM', '') == ''     # A/B switch for the fused ToImage epilogue


def _clp(x, conv, leak, save):
    y, r = ops.conv3x3_fwd(x, conv_images(conv)[0], conv.bias.detach() if conv.bias is not None else None,
                           conv.scale_value, leak, conv.out_channels, want_r=save)
    return y, r


def _clp_toim(x, conv, leak, save, toim, need_y, img_out=None):
    """conv + LeakyReLU + PixelNorm followed by ToImage (`toim`): one kernel when the layer runs on the folded conv
    kernel (its epilogue has every channel of a pixel in registers), else two.  y is not even stored when neither
    the backward pass (save) nor a following block (need_y) wants it.  Returns (y | None, r | None, img)."""
    w_im = _flat(toim.weight)
    if FUSE_TOIMAGE and ops.conv_weight_is_folded(conv.in_channels, conv.out_channels):
        return ops.conv3x3_fwd_toim(x, conv_images(conv)[0], conv.bias.detach() if conv.bias is not None else None,
                                    conv.scale_value, leak, conv.out_channels, w_im, want_y=save or need_y,
                                    want_r=save, img_out=img_out)
    y, r = _clp(x, conv, leak, save)
    return y, r, ops.toim_fwd(y, w_im, out=img_out)


def _g_block_fwd(blk, y, leak, save, toim=None, need_y=True, img_out=None):
    xu = ops.upsample2x(y)
    y1, r1 = _clp(xu, blk.conv1, leak, save)
    if toim is None:
        y2, r2 = _clp(y1, blk.conv2, leak, save)
        img = None
    else:
        y2, r2, img = _clp_toim(y1, blk.conv2, leak, save, toim, need_y, img_out)
    return SimpleNamespace(blk=blk, xu=xu if save else None, y1=y1 if save else None, r1=r1, y2=y2, r2=r2, img=img)


def _alpha_terms(net, alpha):
    """(alpha, 1 - alpha) as kernel scalar arguments.  A caller that replays captured graphs (TrainStep) hangs a
    2-element device tensor [alpha, 1 - alpha] on the network as `_ngan_alpha_dev`; the kernels then read the values
    from it, so one graph serves every epoch of a fade-in (`_ngan_alpha_dev_value` is the alpha the table holds; a
    stale table is ignored).  Otherwise plain floats."""
    tab = getattr(net, '_ngan_alpha_dev', None)
    if tab is None or getattr(net, '_ngan_alpha_dev_value', None) != alpha:     # no table, or not current: floats
        return alpha, 1.0 - alpha
    return (1.0, tab[0:1]), (1.0, tab[1:2])


def _times(scale, term):
    """scale (float, or already carrying a device factor) times an _alpha_terms entry."""
    if isinstance(term, tuple):
        assert not isinstance(scale, tuple), 'alpha enters every gradient path at most once'
        return ops.scalar_mul(scale, term)
    return ops.scalar_mul(term, scale) if isinstance(scale, tuple) else scale * term


def g_forward(net, z, save, img_out=None):
    """Generator_PG.forward (reference models.py:344-353). z: [B, latent] fp32 -> (img [B, R, R] fp32, ctx).
    img_out: optional preallocated [B, R, R] tensor that receives the image: fp32, or bf16 (generator-only
    inference, BASELINE config 5: written straight from the fused ToImage epilogue when the last conv has one)."""
    leak = net.LeakyReLU_neg_slope
    alpha = net.alpha_value()
    lin, conv0 = net.layers[0], net.layers[4]
    bf16_out = None
    if img_out is not None and img_out.dtype == torch.bfloat16:
        last_conv = net.trunk_blocks()[-1].conv2 if net.trunk_blocks() else conv0
        fused = FUSE_TOIMAGE and ops.conv_weight_is_folded(last_conv.in_channels, last_conv.out_channels)
        if save or alpha < 1 or not fused:        # the image passes through fp32 kernels first: convert at the end
            bf16_out, img_out = img_out, None
    S, C0 = net.image_size_init, net.N_features_per_layer[0]
    lin._ngan_dims = (lin.in_features, C0, S)
    z = z.detach().to(F32).contiguous()
    y0, r0 = ops.linear_fwd(z, linear_shadow(lin), lin.scale_value, leak, C0, S, want_r=save)
    fade = alpha < 1
    blocks = net.trunk_blocks()
    # the trunk's ToImage rides on the trunk's last conv; during a fade-in its activation also feeds the new block
    trunk_out = img_out if not fade else None
    if blocks:
        yc, rc = _clp(y0, conv0, leak, save)
        img_trunk = None
    else:
        yc, rc, img_trunk = _clp_toim(y0, conv0, leak, save, net.ToIm, need_y=fade, img_out=trunk_out)
    recs, y, r = [], yc, rc
    for i, blk in enumerate(blocks):
        last = i == len(blocks) - 1
        rec = _g_block_fwd(blk, y, leak, save, toim=net.ToIm if last else None, need_y=fade,
                           img_out=trunk_out if last else None)
        recs.append(rec)
        y, r = rec.y2, rec.r2
        if last:
            img_trunk = rec.img
    ctx = (SimpleNamespace(z=z, y0=y0, r0=r0, yc=yc, rc=rc, recs=recs, alpha=alpha, net=net, new=None, toim=net.ToIm,
                           toim_new=None) if save else None)
    if not fade:
        if save:
            ctx.img = img_trunk
        if bf16_out is not None:
            return ops.f32_to_bf16(img_trunk, bf16_out), ctx
        return img_trunk, ctx
    # fade-in: im_start = up(ToIm_old(x)), im_end = ToIm_new(block_new(x))      (models.py:347-350)
    new = _g_block_fwd(net.conv_block_list[0], y, leak, save, toim=net.ToIm_list[0], need_y=False)
    a_dev, _ = _alpha_terms(net, alpha)
    img = ops.lerp(ops.up2_image(img_trunk), new.img, a_dev, out=img_out)
    if save:
        ctx.new, ctx.img_old, ctx.img_end, ctx.toim_new = new, img_trunk, new.img, net.ToIm_list[0]
    if bf16_out is not None:
        return ops.f32_to_bf16(img, bf16_out), ctx
    return img, ctx


def g_ctx_slice(ctx, lo, hi):
    """The saved generator context of samples [lo, hi) of a forward pass over a larger batch (TrainStep runs the three
    generator passes of an iteration as ONE batch at the low resolutions and backpropagates through the last third)."""
    def cut(v):
        if torch.is_tensor(v) and v.dim() > 0:
            return v[lo:hi]
        if isinstance(v, SimpleNamespace):
            return SimpleNamespace(**{k: cut(x) for k, x in vars(v).items()})
        if isinstance(v, list):
            return [cut(x) for x in v]
        return v
    keep = {k: v for k, v in vars(ctx).items() if k in ('net', 'toim', 'toim_new', 'alpha')}
    out = SimpleNamespace(**{k: cut(v) for k, v in vars(ctx).items() if k not in keep})
    for k, v in keep.items():
        setattr(out, k, v)
    return out


def _g_block_bwd(rec, ga2, y_prev, r_prev, leak, sink, extra_pre=None, extra_w=None):
    """Backward through one generator block given ga2 (gradient at conv2's pre-activation); returns the
    gradient at the pre-activation of the stage below (whose output y_prev was upsampled into this block)."""
    c1, c2 = rec.blk.conv1, rec.blk.conv2
    acc = _acc(sink)
    _wgrad(rec.y1, ga2, c2.scale_value, _sink_get(sink, c2.weight), acc)
    ga1, _ = ops.conv3x3_dgrad_pn(ga2, conv_images(c2)[1], c2.scale_value, leak, rec.y1, rec.r1)
    _wgrad(rec.xu, ga1, c1.scale_value, _sink_get(sink, c1.weight), acc)
    g_up = ops.conv3x3_dgrad(ga1, conv_images(c1)[1], c1.scale_value, c1.in_channels)
    return ops.up2_bwd_pn_bwd(g_up, y_prev, r_prev, extra_pre, extra_w, leak)


def g_backward(net, ctx, g_img, sink, linear_overwrite=False, linear_factors=None):
    """Gradient of the generator parameters given d loss / d image ([B, R, R] fp32).  linear_overwrite: the Linear
    weight's gradient (98 % of the parameters) is stored, not accumulated -- its part of the sink need not be zeroed.
    linear_factors: a SimpleNamespace; when given, the Linear weight's gradient is NOT formed: its two factors (the
    gradient at the stem's pre-activation `ga0`, and the latents `z`) are left on it for ops.adam_linear_factored."""
    leak = net.LeakyReLU_neg_slope
    alpha = ctx.alpha
    acc = _acc(sink)
    g_img = g_img.detach().to(F32).contiguous()
    trunk_y, trunk_r = (ctx.recs[-1].y2, ctx.recs[-1].r2) if ctx.recs else (ctx.yc, ctx.rc)
    if ctx.new is None:
        ga, _ = ops.toim_bwd(g_img, ctx.img, trunk_y, trunk_r, _flat(ctx.toim.weight), _sink_get(sink, ctx.toim.weight),
                             leak=leak, grad_accumulate=acc)
    else:
        new, toim_new, toim_old = ctx.new, ctx.toim_new, ctx.toim
        a_dev, oma_dev = _alpha_terms(net, alpha)
        ga2, _ = ops.toim_bwd(g_img, ctx.img_end, new.y2, new.r2, _flat(toim_new.weight),
                              _sink_get(sink, toim_new.weight), gscale=a_dev, leak=leak, grad_accumulate=acc)
        g_old = ops.up2_image_bwd(g_img, oma_dev)
        _, gpre_old = ops.toim_bwd(g_old, ctx.img_old, trunk_y, None, _flat(toim_old.weight),
                                   _sink_get(sink, toim_old.weight), want_ga=False, want_gpre=True, leak=leak,
                                   grad_accumulate=acc)
        ga = _g_block_bwd(new, ga2, trunk_y, trunk_r, leak, sink, gpre_old, _flat(toim_old.weight))
    for i in range(len(ctx.recs) - 1, -1, -1):
        y_prev, r_prev = (ctx.recs[i - 1].y2, ctx.recs[i - 1].r2) if i > 0 else (ctx.yc, ctx.rc)
        ga = _g_block_bwd(ctx.recs[i], ga, y_prev, r_prev, leak, sink)
    lin, conv0 = net.layers[0], net.layers[4]
    _wgrad(ctx.y0, ga, conv0.scale_value, _sink_get(sink, conv0.weight), acc)
    ga0, _ = ops.conv3x3_dgrad_pn(ga, conv_images(conv0)[1], conv0.scale_value, leak, ctx.y0, ctx.r0)
    if linear_factors is not None:
        linear_factors.ga0, linear_factors.z, linear_factors.scale = ga0, ctx.z, lin.scale_value
        return
    ops.linear_wgrad(ga0, ctx.z, lin.scale_value, _sink_get(sink, lin.weight), accumulate=acc and not linear_overwrite)


# ---------------------------------------------------------------------------------------------------------
# Discriminator
# ---------------------------------------------------------------------------------------------------------
def d_forward(net, x, save):
    """Discriminator_PG.forward (reference models.py:516-524). x: [B, R, R] fp32 -> (scores [B] fp32, ctx).

    AvgPool2d(2) commutes with the 1x1 FromImage conv and the fade-in branch's bilinear x0.5 is the same 2x2
    mean, so both branches start from one pooled 1-channel image `xp` and the
    full-resolution F-channel tensor is never materialised."""
    leak = net.LeakyReLU_neg_slope
    alpha = net.alpha_value()
    x = x.detach().to(F32).contiguous()
    trunk = net.trunk_blocks()
    pooled = net.N_layers > 1
    xp = ops.pool_image(x) if pooled else x
    stages = []

    def block(blk, xin, pooled_in):
        y1, r1 = _clp(xin, blk.conv1, leak, save)
        y2, r2 = _clp(y1, blk.conv2, leak, save)
        return SimpleNamespace(kind='block', blk=blk, xin=xin if save else None, xin_pooled=pooled_in,
                               y1=y1 if save else None, r1=r1, y2=y2, r2=r2, out=y2, consumer_pooled=False)

    if alpha < 1:
        f_new = net.FromIm_list[-1].conv
        F = ops.fromim_fwd(xp, _flat(f_new.weight), f_new.bias.detach())
        stages.append(SimpleNamespace(kind='from', conv=f_new, out=F, consumer_pooled=False))
        stages.append(block(net.conv_block_list[-1], F, False))
        f_old = net.FromIm.conv
        y = ops.d_fade_fwd(stages[-1].out, xp, _flat(f_old.weight), f_old.bias.detach(), _alpha_terms(net, alpha)[0])
        stages.append(SimpleNamespace(kind='fade', conv=f_old, out=y, consumer_pooled=False))
        rest = trunk
    else:
        f_cur = net.FromIm.conv
        F = ops.fromim_fwd(xp, _flat(f_cur.weight), f_cur.bias.detach())
        stages.append(SimpleNamespace(kind='from', conv=f_cur, out=F, consumer_pooled=False))
        rest = trunk
        if trunk:  # the first trunk block's AvgPool is already folded into xp
            stages.append(block(trunk[0], F, False))
            rest = trunk[1:]
    for blk in rest:
        stages[-1].consumer_pooled = True
        stages.append(block(blk, ops.avgpool2(stages[-1].out), True))
    last_conv, head = net.last_conv(), net.head_conv()
    yl, rl = _clp(stages[-1].out, last_conv, leak, save)
    scores = ops.head_fwd(yl, head.weight.detach(), head.bias.detach(), head.scale_value)
    if not save:
        return scores, None
    ctx = SimpleNamespace(xp=xp, pooled=pooled, alpha=alpha, stages=stages, last_x=stages[-1].out, yl=yl, rl=rl,
                          B=x.shape[0])
    return scores, ctx


def d_backward(net, ctx, gout, sink, addins=None, want_gxp=False, record=None):
    """Backward through the critic.

    gout   : d loss / d score [B] fp32, or None for sweep 2 of the double backward (no head contribution).
    sink   : parameter-gradient accumulators, or None to skip every weight/bias gradient (input gradient only).
    addins : per-stage cotangents injected at the pre-activations (from d_double_backward_sweep1).
    record : SimpleNamespace to fill with the first-order gradients the double backward needs.
    Returns g_xp, the gradient wrt the pooled image `xp` ([B, h, w] fp32) when want_gxp (callers un-pool it)."""
    leak = net.LeakyReLU_neg_slope
    alpha = ctx.alpha
    a_dev, oma_dev = _alpha_terms(net, alpha)
    last, head = net.last_conv(), net.head_conv()
    rec = record is not None
    ad = addins or {}
    acc = _acc(sink)
    if gout is not None:
        gout = gout.detach().to(F32).contiguous()
        ga_l, gy_l = ops.head_bwd_pn(gout, head.weight.detach(), head.scale_value, ctx.yl, ctx.rl, want_gy=rec,
                                     leak=leak)
        if sink is not None:
            hw, hb, yl = _sink_get(sink, head.weight), _sink_get(sink, head.bias), ctx.yl
            _on_side(lambda: ops.head_wgrad(yl, gout, head.scale_value, hw, hb, accumulate=acc), yl, gout)
    else:
        ga_l, gy_l = ad['last'], None
    if rec:
        record.last = SimpleNamespace(gy=gy_l, ga=ga_l)
        record.gout = gout
        record.stages = {}
    if sink is not None:
        _wgrad(ctx.last_x, ga_l, last.scale_value, _sink_get(sink, last.weight), acc)
        lb = _sink_get(sink, last.bias)
        _on_side(lambda: ops.bias_grad(ga_l, lb, accumulate=acc), ga_l)

    top = ctx.stages[-1]
    if top.kind == 'block':
        a2 = ad.get(id(top), (None, None))[1]
        ga2, gy2 = ops.conv3x3_dgrad_pn(ga_l, conv_images(last)[1], last.scale_value, leak, top.y2, top.r2, addin=a2,
                                        want_gy=rec)
        incoming = ('ga', ga2, gy2)
    else:
        incoming = ('g', ops.conv3x3_dgrad(ga_l, conv_images(last)[1], last.scale_value, last.in_channels), 1.0,
                    False)

    # `incoming` is the gradient handed to the stage being visited: either ('ga', ga2, gy2) -- already through the
    # PixelNorm/LeakyReLU backward -- or ('g', tensor, scale, unpool): scale * (AvgPool adjoint if unpool)(tensor)
    # is the gradient wrt the stage's output.
    g_xp = None
    for st in reversed(ctx.stages):
        if st.kind == 'block':
            c1, c2 = st.blk.conv1, st.blk.conv2
            a1, a2 = ad.get(id(st), (None, None))
            if incoming[0] == 'ga':
                _, ga2, gy2 = incoming
                recv_unpool = False
            else:
                _, g, gs, recv_unpool = incoming
                ga2, gy2 = ops.pn_bwd(g, st.y2, st.r2, gscale=ops.scalar_mul(0.25 if recv_unpool else 1.0, gs),
                                      unpool=recv_unpool, addin=a2, want_gy=rec, leak=leak)
            if sink is not None:
                _wgrad(st.y1, ga2, c2.scale_value, _sink_get(sink, c2.weight), acc)
            ga1, gy1 = ops.conv3x3_dgrad_pn(ga2, conv_images(c2)[1], c2.scale_value, leak, st.y1, st.r1, addin=a1,
                                            want_gy=rec)
            if sink is not None:
                _wgrad(st.xin, ga1, c1.scale_value, _sink_get(sink, c1.weight), acc)
            if rec:
                record.stages[id(st)] = SimpleNamespace(gy1=gy1, ga1=ga1, gy2=gy2, ga2=ga2, recv_unpool=recv_unpool)
            incoming = ('g', ops.conv3x3_dgrad(ga1, conv_images(c1)[1], c1.scale_value, c1.in_channels), 1.0,
                        st.xin_pooled)
        else:
            _, g, gs, unpool = incoming
            gs_eff = ops.scalar_mul(0.25 if unpool else 1.0, gs)
            if rec:
                record.stages[id(st)] = SimpleNamespace(g=g, unpool=unpool, gscale=gs_eff)
            need_img = want_gxp
            if need_img and g_xp is None:
                g_xp = torch.empty_like(ctx.xp)
                accumulate = False
            else:
                accumulate = True
            w = _flat(st.conv.weight)
            if st.kind == 'fade':
                if sink is not None or need_img:
                    ops.fromim_bwd(g, ctx.xp, w, _sink_get(sink, st.conv.weight), _sink_get(sink, st.conv.bias),
                                   gscale=_times(gs_eff, oma_dev), unpool=unpool, g_img=g_xp if need_img else None,
                                   accumulate=accumulate, grad_accumulate=acc)
                incoming = ('g', g, _times(gs, a_dev), unpool)
            else:  # 'from'
                if sink is not None or need_img:
                    ops.fromim_bwd(g, ctx.xp, w, _sink_get(sink, st.conv.weight), _sink_get(sink, st.conv.bias),
                                   gscale=gs_eff, unpool=unpool, g_img=g_xp if need_img else None,
                                   accumulate=accumulate, grad_accumulate=acc)
    return g_xp


def d_double_backward_sweep1(net, ctx, record, ghat_xp, sink):
    """Sweep 1 of the gradient-penalty double backward: propagate the cotangent on the first-order input
    gradient (ghat_xp, [B, h, w] fp32, living on the pooled image) from the image side up to the head.
    Accumulates the "wgrad of dgrad" terms into `sink` and returns the per-stage injections for sweep 2."""
    leak = net.LeakyReLU_neg_slope
    alpha = ctx.alpha
    a_dev, oma_dev = _alpha_terms(net, alpha)
    addins = {}
    cur = None
    acc = _acc(sink)
    ghat_xp = ghat_xp.contiguous()
    for st in ctx.stages:
        r = record.stages[id(st)]
        if st.kind == 'from':
            cur = ops.fromim_dbl(ghat_xp, r.g, _flat(st.conv.weight), _sink_get(sink, st.conv.weight), in_scale=1.0,
                                 gscale=r.gscale, unpool=r.unpool, grad_accumulate=acc)
            if r.unpool:            # cannot happen with the reference's topology, kept for completeness
                cur = ops.avgpool2(cur)
        elif st.kind == 'block':
            c1, c2 = st.blk.conv1, st.blk.conv2
            _wgrad(cur, r.ga1, c1.scale_value, _sink_get(sink, c1.weight), acc)
            gh1, ah1 = ops.conv3x3_dbl(cur, conv_images(c1)[0], c1.scale_value, leak, st.y1, st.r1, r.gy1)
            _wgrad(gh1, r.ga2, c2.scale_value, _sink_get(sink, c2.weight), acc)
            gh2, ah2 = ops.conv3x3_dbl(gh1, conv_images(c2)[0], c2.scale_value, leak, st.y2, st.r2, r.gy2)
            addins[id(st)] = (ah1, ah2)
            cur = gh2          # cotangent on gy2; the fade stage (if next) applies alpha and the pooling itself
            nxt = ctx.stages.index(st) + 1
            to_fade = nxt < len(ctx.stages) and ctx.stages[nxt].kind == 'fade'
            if r.recv_unpool and not to_fade:
                cur = ops.avgpool2(cur)
        else:  # 'fade': first-order g_y reached FromIm_old with (1-alpha) and the new block with alpha
            w_old = _flat(st.conv.weight)
            ops.fromim_dbl(ghat_xp, r.g, w_old, _sink_get(sink, st.conv.weight), in_scale=1.0,
                           gscale=_times(r.gscale, oma_dev), unpool=r.unpool, want_out=False, grad_accumulate=acc)
            cur = ops.d_fade_fwd(cur, ghat_xp, w_old, None, a_dev)     # alpha*cur + (1-alpha)*w_old*ghat_xp
            if r.unpool:
                cur = ops.avgpool2(cur)
    last, head = net.last_conv(), net.head_conv()
    _wgrad(cur, record.last.ga, last.scale_value, _sink_get(sink, last.weight), acc)
    gh_l, ah_l = ops.conv3x3_dbl(cur, conv_images(last)[0], last.scale_value, leak, ctx.yl, ctx.rl, record.last.gy)
    hw, rgout = _sink_get(sink, head.weight), record.gout
    _on_side(lambda: ops.head_wgrad(gh_l, rgout, head.scale_value, hw, accumulate=acc), gh_l, rgout)
    addins['last'] = ah_l
    return addins


_ones_cache = {}


def _ones(n, device):
    """[n] fp32 ones, created once per (n, device): no fill kernel inside a captured iteration."""
    key = (n, str(device))
    t = _ones_cache.get(key)
    if t is None:
        t = _ones_cache[key] = torch.ones(n, dtype=F32, device=device)
    return t


def d_grad_penalty(net, x_hat, lam, sink, gscale=1.0):
    """The whole gradient-penalty term (reference loss_functions.py:157-180) for an already interpolated batch:
    forward, first-order input gradient, penalty, and -- when `sink` is given -- its parameter gradients
    (scaled by gscale).  Returns (penalty [1] fp32, callable that runs the double backward into a sink)."""
    scores, ctx = d_forward(net, x_hat, save=True)
    B = ctx.B
    ones = _ones(B, scores.device)
    record = SimpleNamespace()
    g_xp = d_backward(net, ctx, ones, None, want_gxp=True, record=record)
    ns = 0.5 if ctx.pooled else 1.0       # ||unpool(g)/4|| = ||g|| / 2
    pen, coeff = ops.gp_loss(g_xp, ns, float(lam), gscale=1.0)

    def backward(into, scale=1.0, scale_tensor=None):
        """into: one sink for both sweeps (its kernels then run in stream order on ONE stream -- see below), or a pair
        (sink of sweep 1, sink of sweep 2): the two sweeps contribute to the same parameters from side streams."""
        c = coeff if scale_tensor is None else coeff * scale_tensor
        ghat_xp = ops.scale_rows(g_xp, c, ns * ns * scale)
        s1, s2 = into if isinstance(into, tuple) else (into, into)
        addins = d_double_backward_sweep1(net, ctx, record, ghat_xp, s1)
        if s1 is s2 and s1 is not None:
            side_join()        # sweep 2 read-modify-writes what sweep 1's side-stream kernels are still producing
        d_backward(net, ctx, None, s2, addins=addins)

    if sink is not None:
        backward(sink, gscale)
    return pen, backward, (g_xp, ctx)
