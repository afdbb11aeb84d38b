"""Debug aid (GPU box): stage-by-stage comparison of the critic's forward activations and first-order
backward intermediates against the bf16-emulating oracle, for gout = given per-sample coefficients."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import pggan_oracle as O
from neuron_gan_b200 import engine, ops
from neuron_gan_b200.train_step import build_networks
from types import SimpleNamespace

ARCH = O.Arch()
DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def oracle_d(p, x, n_layers, alpha):
    """d_forward of the oracle, re-stated here to capture intermediates (bf16 emulation on)."""
    leak, L = ARCH.leak, len(ARCH.dis_features)
    top = L - n_layers
    caps = {}

    def clp(x, w, b, tag):
        a = O.eq_conv(x, w, b, 1, leak)
        a.retain_grad()
        y = O._q(O.pixel_norm(F.leaky_relu(a, leak)))
        y.retain_grad()
        caps[tag] = (a, y)
        return y

    def block(x, i, pool):
        if pool:
            x = O._q(F.avg_pool2d(x, 2))
        x.retain_grad()
        caps[f'blk{i}.xin'] = (x, x)
        return clp(clp(x, p[f'blk{i}.w1'], None, f'blk{i}.1'), p[f'blk{i}.w2'], None, f'blk{i}.2')

    if alpha >= 1:
        y = O.from_image(x, p[f'from{top}.w'], p[f'from{top}.b'])
        first = top
        if first == L - 1:
            y = O._q(y)
    else:
        ys = O.from_image(O.down2_bilinear(x), p[f'from{top + 1}.w'], p[f'from{top + 1}.b'])
        ye = block(O.from_image(x, p[f'from{top}.w'], p[f'from{top}.b']), top, True)
        y = O._q(ys + alpha * (ye - ys))
        y.retain_grad()
        caps['fade'] = (y, y)
        first = top + 1
    for i in range(first, L - 1):
        y = block(y, i, True)
    y = clp(y, p['last.w'], p['last.b'], 'last')
    out = O.eq_conv(y, p['head.w'], p['head.b'], 0, leak).flatten(1)
    return out, caps


def main(res, alpha, B, mode):
    G, D = build_networks(res, alpha, seed=1, device=DEV)
    n = O.n_layers_for(res, ARCH)
    L = len(ARCH.dis_features)
    dkm = O.d_key_map(n, alpha < 1, ARCH)
    ds = D.state_dict()
    dp = {k: ds[v].detach().cpu().clone().requires_grad_() for k, v in dkm.items()}
    x = O.synthetic_images(B, res, seed=13)
    torch.manual_seed(3)
    coef = torch.randn(B) if mode == 'rand' else torch.ones(B)
    with O.emulate_bf16():
        out, caps = oracle_d(dp, x, n, alpha)
        (out[:, 0] * coef).sum().backward()
    with torch.no_grad():
        scores, ctx = engine.d_forward(D, x[:, 0].to(DEV), save=True)
        rec = SimpleNamespace()
        sink = {id(p): torch.zeros_like(p) for p in D.active_parameters()}
        engine.d_backward(D, ctx, coef.to(DEV), sink, want_gxp=True, record=rec)
    print(f'--- res={res} alpha={alpha} B={B} coef={mode}: scores rel {rel(scores.cpu(), out[:, 0].detach()):.5f}')
    top = L - n
    blocks = [s for s in ctx.stages if s.kind == 'block']
    idx = top
    for st in blocks:
        r = rec.stages[id(st)]
        a1, y1 = caps[f'blk{idx}.1']
        a2, y2 = caps[f'blk{idx}.2']
        xin = caps[f'blk{idx}.xin'][0]
        print(f' blk{idx}: xin {rel(ops.c8_to_nchw(st.xin).cpu(), xin.detach()):.4f} y1 {rel(ops.c8_to_nchw(st.y1).cpu(), y1.detach()):.4f} '
              f'y2 {rel(ops.c8_to_nchw(st.y2).cpu(), y2.detach()):.4f} | gy2 {rel(ops.c8_to_nchw(r.gy2).cpu(), y2.grad):.4f} '
              f'ga2 {rel(ops.c8_to_nchw(r.ga2).cpu(), a2.grad):.4f} gy1 {rel(ops.c8_to_nchw(r.gy1).cpu(), y1.grad):.4f} '
              f'ga1 {rel(ops.c8_to_nchw(r.ga1).cpu(), a1.grad):.4f}')
        idx += 1
    al, yl = caps['last']
    print(f' last: y {rel(ops.c8_to_nchw(ctx.yl).cpu(), yl.detach()):.4f} gy {rel(ops.c8_to_nchw(rec.last.gy).cpu(), yl.grad):.4f} '
          f'ga {rel(ops.c8_to_nchw(rec.last.ga).cpu(), al.grad):.4f}')
    if 'fade' in caps:
        st = [s for s in ctx.stages if s.kind == 'fade'][0]
        print(f' fade: y {rel(ops.c8_to_nchw(st.out).cpu(), caps["fade"][0].detach()):.4f}')
    named = dict(D.named_parameters())
    print(' param grads: ' + ' '.join(f'{k}:{rel(sink[id(named[v])].cpu(), dp[k].grad):.3f}' for k, v in dkm.items()
                                       if dp[k].grad is not None and id(named[v]) in sink))


if __name__ == '__main__':
    for a in sys.argv[1:]:
        res, alpha, B, mode = a.split(',')
        main(int(res), float(alpha), int(B), mode)
