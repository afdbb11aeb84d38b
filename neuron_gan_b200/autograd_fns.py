"""autograd glue: the networks appear to PyTorch as two opaque, differentiable Functions each, so the
reference's own call patterns keep working on these modules --

    loss.backward()                                                   (train.py:365, 384)
    torch.autograd.grad(D(x_hat).sum(), x_hat, create_graph=True)     (loss_functions.py:175)

The critic's backward is itself a Function whose backward is the hand-written double backward.  Everything
inside the Functions is neuron_gan_b200.engine (explicit kernel sequences); autograd only routes cotangents.
"""
from types import SimpleNamespace

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import engine, ops

F32 = torch.float32


def _zeros_sink(params):
    return {id(p): torch.zeros(p.shape, dtype=F32, device=p.device) for p in params}


# ------------------------------------------------------------------------------------------------ generator
class _GeneratorFn(Function):
    @staticmethod
    def forward(ctx, net, plist, z, *params):
        img, ectx = engine.g_forward(net, z, save=True)
        ctx.net, ctx.ectx, ctx.plist = net, ectx, plist
        return img.unsqueeze(1)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_img):
        sink = _zeros_sink(ctx.plist)
        engine.g_backward(ctx.net, ctx.ectx, g_img[:, 0].contiguous(), sink)
        engine.side_join()
        ctx.ectx = None
        return (None, None, None) + tuple(sink[id(p)] for p in ctx.plist)


def generator_forward(net, z):
    plist = net.active_parameters()
    if torch.is_grad_enabled() and any(p.requires_grad for p in plist):
        return _GeneratorFn.apply(net, plist, z, *plist)
    img, _ = engine.g_forward(net, z, save=False)
    return img.unsqueeze(1)


# ------------------------------------------------------------------------------------------------ critic
class _CriticBackwardFn(Function):
    """First-order backward of the critic as a differentiable op: (gout, params) -> (gx, param grads)."""

    @staticmethod
    def forward(ctx, holder, want_gx, want_pgrads, record, gout, *params):
        net, ectx, plist = holder.net, holder.ectx, holder.plist
        rec = SimpleNamespace() if record else None
        sink = _zeros_sink(plist) if want_pgrads else None
        g_xp = engine.d_backward(net, ectx, gout[:, 0].contiguous(), sink, want_gxp=want_gx, record=rec)
        engine.side_join()
        ctx.holder, ctx.rec = holder, rec
        ctx.set_materialize_grads(False)
        gx = None
        if want_gx:
            gx = (ops.unpool_image(g_xp, 0.25) if ectx.pooled else g_xp).unsqueeze(1)
        pg = tuple(sink[id(p)] if sink is not None else None for p in plist)
        return (gx,) + pg

    @staticmethod
    @once_differentiable
    def backward(ctx, ghat_x, *ghat_params):
        holder = ctx.holder
        net, ectx, plist = holder.net, holder.ectx, holder.plist
        if any(g is not None for g in ghat_params):
            raise NotImplementedError('differentiating through the critic parameter gradients is not supported')
        if ghat_x is None:
            return (None,) * (5 + len(plist))
        if ctx.rec is None:
            raise RuntimeError('double backward requested but the first backward ran without create_graph=True')
        gh = ghat_x[:, 0].contiguous().to(F32)
        ghat_xp = ops.pool_image(gh) if ectx.pooled else gh
        sink = _zeros_sink(plist)
        addins = engine.d_double_backward_sweep1(net, ectx, ctx.rec, ghat_xp, sink)
        engine.side_join()      # sweep 2 adds to the tensors sweep 1's side-stream kernels are still writing
        engine.d_backward(net, ectx, None, sink, addins=addins)
        engine.side_join()
        return (None, None, None, None, None) + tuple(sink[id(p)] for p in plist)


class _CriticFn(Function):
    @staticmethod
    def forward(ctx, net, plist, params_grad, x, *params):
        scores, ectx = engine.d_forward(net, x[:, 0], save=True)
        ctx.holder = SimpleNamespace(net=net, ectx=ectx, plist=plist)
        ctx.params_grad = params_grad
        ctx.params = params
        return scores.unsqueeze(1)

    @staticmethod
    def backward(ctx, gout):
        want_gx = ctx.needs_input_grad[3]
        want_pg = ctx.params_grad and any(ctx.needs_input_grad[4:])
        outs = _CriticBackwardFn.apply(ctx.holder, want_gx, want_pg, torch.is_grad_enabled(), gout.contiguous(),
                                       *ctx.params)
        gx, pg = outs[0], outs[1:]
        if not want_pg:
            pg = (None,) * len(ctx.params)
        return (None, None, None, gx) + tuple(pg)


def discriminator_forward(net, x, params_grad=True):
    """params_grad=False treats the critic's parameters as constants (generator step: only d/dx is needed)."""
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[-1] != net.image_size or x.shape[-2] != net.image_size:
        raise ValueError(f'expected images of shape [B, 1, {net.image_size}, {net.image_size}], got {tuple(x.shape)}')
    plist = net.active_parameters()
    needs = torch.is_grad_enabled() and (x.requires_grad or (params_grad and any(p.requires_grad for p in plist)))
    if needs:
        return _CriticFn.apply(net, plist, params_grad, x.to(F32), *plist)
    scores, _ = engine.d_forward(net, x[:, 0], save=False)
    return scores.unsqueeze(1)


# ------------------------------------------------------------------------------------------------ losses
class _WLossFn(Function):
    """D_W_loss arithmetic (reference loss_functions.py:20-45) on the scores of one [real; fake] critic pass."""

    @staticmethod
    def forward(ctx, scores, B, drift):
        s = scores.reshape(-1).contiguous()
        out3, g_real, g_fake = ops.wloss(s[:B], s[B:], float(drift))
        ctx.g = torch.cat([g_real, g_fake]).unsqueeze(1)
        ctx.B = B
        ctx.set_materialize_grads(False)
        return out3[0].clone(), out3[1].clone(), out3[2].clone()   # callers do `D_loss += pen` in place

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loss, g_sr, g_sf):
        g = ctx.g * g_loss if g_loss is not None else torch.zeros_like(ctx.g)
        if g_sr is not None:
            g[:ctx.B] += g_sr / ctx.B
        if g_sf is not None:
            g[ctx.B:] += g_sf / ctx.B
        return g, None, None


class _GLossFn(Function):
    """G_W_loss arithmetic (reference loss_functions.py:67): -mean(D(G(z)))."""

    @staticmethod
    def forward(ctx, scores):
        out1, g_fake = ops.gloss(scores.reshape(-1).contiguous())
        ctx.g = g_fake.unsqueeze(1)
        return out1[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_loss):
        return ctx.g * g_loss


class _GradPenaltyFn(Function):
    """The fused gradient-penalty term (reference loss_functions.py:173-176) for an interpolated batch."""

    @staticmethod
    def forward(ctx, net, plist, lam, x_hat, *params):
        pen, closure, _ = engine.d_grad_penalty(net, x_hat[:, 0], lam, sink=None)
        ctx.closure, ctx.plist = closure, plist
        return pen[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_pen):
        sink = _zeros_sink(ctx.plist)
        ctx.closure(sink, 1.0, scale_tensor=g_pen)
        engine.side_join()
        ctx.closure = None
        return (None, None, None, None) + tuple(sink[id(p)] for p in ctx.plist)


def gradient_penalty(net, x_hat, lam):
    plist = net.active_parameters()
    return _GradPenaltyFn.apply(net, plist, float(lam), x_hat.to(F32), *plist)
