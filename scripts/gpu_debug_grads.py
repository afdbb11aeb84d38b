"""Debug aid (GPU box): split the gradient comparison against the bf16-emulating oracle by loss term."""
import sys
import torch
sys.path.insert(0, '.')
from oracle import pggan_oracle as O
from neuron_gan_b200 import engine, ops
from neuron_gan_b200.train_step import build_networks

ARCH = O.Arch()
DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main(res, alpha, B):
    G, D = build_networks(res, alpha, seed=1, device=DEV)
    n = O.n_layers_for(res, ARCH)
    gkm, dkm = O.g_key_map(n, alpha < 1, ARCH), O.d_key_map(n, alpha < 1, ARCH)
    gs, ds = G.state_dict(), D.state_dict()
    gp = {k: gs[v].detach().cpu().clone().requires_grad_() for k, v in gkm.items()}
    dp = {k: ds[v].detach().cpu().clone().requires_grad_() for k, v in dkm.items()}
    x = O.synthetic_images(B, res, seed=13)
    torch.manual_seed(3)
    z1, z2, z3 = (O.sample_latent((B, 512)) for _ in range(3))
    eps = torch.rand(B, 1, 1, 1)
    dn, gn = O.active_d_names(n, alpha, ARCH), O.active_g_names(n, alpha, ARCH)
    dnamed, gnamed = dict(D.named_parameters()), dict(G.named_parameters())

    def sink_for(net):
        return {id(p): torch.zeros_like(p) for p in net.active_parameters()}

    def report(tag, sink, named, km, names, ref):
        out = {}
        for k in names:
            if ref[k] is None:
                continue
            out[k] = rel(sink[id(named[km[k]])].cpu(), ref[k])
        v = sorted(out.values())
        print(f'[{tag}] res={res} a={alpha}: median {v[len(v)//2]:.4f} max {v[-1]:.4f}  ' +
              ' '.join(f'{k}:{e:.3f}' for k, e in out.items()), flush=True)

    with torch.no_grad():
        # (1) W loss only
        with O.emulate_bf16():
            pass
    with O.emulate_bf16():
        loss, sr, sf = O.d_w_loss(gp, dp, x, z1, n, alpha, ARCH, drift=1e-3)
        ref = dict(zip(dn, torch.autograd.grad(loss, [dp[k] for k in dn], allow_unused=True)))
    with torch.no_grad():
        fake, _ = engine.g_forward(G, z1.to(DEV), save=False)
        xx = torch.cat([x[:, 0].to(DEV), fake])
        scores, ctx = engine.d_forward(D, xx, save=True)
        gout = torch.empty(2 * B, device=DEV)
        out3, _, _ = ops.wloss_into(scores[:B], scores[B:], 1e-3, gout[:B], gout[B:])
        sink = sink_for(D)
        engine.d_backward(D, ctx, gout, sink)
    print('wloss', out3.tolist(), loss.item(), sr.item(), sf.item())
    report('W-loss D grads', sink, dnamed, dkm, dn, ref)

    # (2) GP only
    with O.emulate_bf16():
        pen, g1 = O.grad_penalty(gp, dp, x, z2, eps, n, alpha, ARCH, lam=10.0, return_grad=True)
        ref = dict(zip(dn, torch.autograd.grad(pen, [dp[k] for k in dn], allow_unused=True)))
    with torch.no_grad():
        xt, _ = engine.g_forward(G, z2.to(DEV), save=False)
        x_hat = ops.interp_images(x[:, 0].to(DEV).contiguous(), xt, eps.reshape(B).to(DEV))
        sink = sink_for(D)
        pen_g, _, (g_xp, dctx) = engine.d_grad_penalty(D, x_hat, 10.0, sink)
        gx = ops.unpool_image(g_xp, 0.25) if dctx.pooled else g_xp
    print('pen', pen_g.item(), pen.item(), 'first-order input grad rel', rel(gx.cpu(), g1[:, 0].detach()))
    report('GP D grads', sink, dnamed, dkm, dn, ref)

    # (3) G step
    with O.emulate_bf16():
        gl = O.g_w_loss(gp, dp, z3, n, alpha, ARCH)
        ref = dict(zip(gn, torch.autograd.grad(gl, [gp[k] for k in gn], allow_unused=True)))
    with torch.no_grad():
        fake, gctx = engine.g_forward(G, z3.to(DEV), save=True)
        s, dctx = engine.d_forward(D, fake, save=True)
        out1, gf = ops.gloss(s)
        g_xp = engine.d_backward(D, dctx, gf, None, want_gxp=True)
        gx = ops.unpool_image(g_xp, 0.25) if dctx.pooled else g_xp
        sink = sink_for(G)
        engine.g_backward(G, gctx, gx, sink)
    print('gloss', out1.item(), gl.item())
    report('G grads', sink, gnamed, gkm, gn, ref)


if __name__ == '__main__':
    cfgs = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(16, 1.0, 8), (32, 0.5, 4), (64, 1.0, 4)]
    for cfg in cfgs:
        main(*cfg)
