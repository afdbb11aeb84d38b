"""Whole-network backward passes on the B200 driven by WELL-CONDITIONED cotangents, against autograd through the
oracle with the CUDA path's storage precision emulated (bf16 weights / feature maps), same weights, same inputs.

Why this file exists: at random initialisation the Wasserstein loss gradient is a difference of nearly equal
real / fake sums, so end-to-end gradient comparisons at 256x256 / 512x512 are limited by conditioning, not by the
kernels (tests/test_step_gpu.py::test_gradients_match_bf16_emulating_oracle carries loose bounds there for that
reason).  Here the cotangents are fixed, same-signed per-sample weights on the critic's scores and a fixed random
image on the generator's output, so nothing cancels and a composition error of a few percent in ONE layer of the
hand-written D backward, G backward or gradient-penalty double backward fails the test.

Tolerance: relative L2 per parameter <= 5e-2; measured values are printed."""
import pytest
import torch

from oracle import pggan_oracle as O

pytestmark = pytest.mark.gpu
ARCH = O.Arch()
DEV = 'cuda'
TOL = 5e-2


def nets(res, alpha):
    from neuron_gan_b200.train_step import build_networks
    return build_networks(res, alpha, seed=1, device=DEV)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def named(net, km):
    sd = net.state_dict()
    return {k: sd[v].detach().cpu().clone().requires_grad_() for k, v in km.items()}


def grads_of(net, km, sink):
    params = dict(net.named_parameters())
    return {name: sink[id(params[key])] for name, key in km.items() if id(params[key]) in sink}


def report_and_check(tag, got, ref, names):
    rep = {}
    for k in names:
        assert ref[k] is not None, k
        rep[k] = rel(got[k].cpu(), ref[k])
    vals = sorted(rep.values())
    worst = max(rep, key=rep.get)
    print(f'{tag}: rel-L2 per parameter median {vals[len(vals) // 2]:.4f} max {vals[-1]:.4f} ({worst})')
    assert vals[-1] <= TOL, rep


CASES = [(64, 0.5, 4), (256, 1.0, 2), (512, 0.5, 1), (512, 1.0, 2)]


@pytest.mark.parametrize('res,alpha,batch', CASES)
def test_critic_backward(res, alpha, batch):
    """d_forward + d_backward with per-sample score cotangents (1.0, 0.8, ...): every critic parameter gradient."""
    from neuron_gan_b200 import autograd_fns, engine
    _, D = nets(res, alpha)
    n = O.n_layers_for(res, ARCH)
    km = O.d_key_map(n, alpha < 1, ARCH)
    dp = named(D, km)
    names = O.active_d_names(n, alpha, ARCH)
    x = O.synthetic_images(batch, res, seed=21)
    gout = torch.tensor([1.0 - 0.2 * (i % 3) for i in range(batch)])
    with O.emulate_bf16():
        scores = O.d_forward(dp, x, n, alpha, ARCH)
        ref = dict(zip(names, torch.autograd.grad((scores[:, 0] * gout).sum(), [dp[k] for k in names])))
    with torch.no_grad():
        s, ctx = engine.d_forward(D, x.to(DEV)[:, 0], save=True)
        sink = autograd_fns._zeros_sink(D.active_parameters())
        engine.d_backward(D, ctx, gout.to(DEV), sink)
        engine.side_join()
    assert torch.allclose(s.cpu(), scores[:, 0].detach(), rtol=0, atol=2e-3)
    report_and_check(f'critic backward r{res} a{alpha}', grads_of(D, km, sink), ref, names)


@pytest.mark.parametrize('res,alpha,batch', CASES)
def test_generator_backward(res, alpha, batch):
    """g_forward + g_backward with a fixed random image cotangent: every generator parameter gradient."""
    from neuron_gan_b200 import autograd_fns, engine
    G, _ = nets(res, alpha)
    n = O.n_layers_for(res, ARCH)
    km = O.g_key_map(n, alpha < 1, ARCH)
    gp = named(G, km)
    names = O.active_g_names(n, alpha, ARCH)
    gen = torch.Generator().manual_seed(5)
    z = O.sample_latent((batch, 512), gen)
    g_img = torch.randn(batch, res, res, generator=gen)
    with O.emulate_bf16():
        img = O.g_forward(gp, z, n, alpha, ARCH)
        ref = dict(zip(names, torch.autograd.grad((img[:, 0] * g_img).sum(), [gp[k] for k in names])))
    with torch.no_grad():
        out, ctx = engine.g_forward(G, z.to(DEV), save=True)
        sink = autograd_fns._zeros_sink(G.active_parameters())
        engine.g_backward(G, ctx, g_img.to(DEV), sink)
        engine.side_join()
    assert rel(out.cpu(), img[:, 0].detach()) < 2e-2
    report_and_check(f'generator backward r{res} a{alpha}', grads_of(G, km, sink), ref, names)


@pytest.mark.parametrize('res,alpha,batch', CASES)
def test_gradient_penalty_double_backward(res, alpha, batch):
    """The gradient penalty (loss_functions.py:173-176) on a fixed batch: penalty value and every critic parameter
    gradient of it -- forward, first-order backward, both sweeps of the hand-written double backward.  All samples
    have ||grad|| << 1 at initialisation, so the per-sample terms have one sign and nothing cancels."""
    from neuron_gan_b200 import autograd_fns
    _, D = nets(res, alpha)
    n = O.n_layers_for(res, ARCH)
    km = O.d_key_map(n, alpha < 1, ARCH)
    dp = named(D, km)
    names = O.active_d_names(n, alpha, ARCH)          # (head.b gets no penalty gradient: autograd returns None)
    x_hat = O.synthetic_images(batch, res, seed=23)
    with O.emulate_bf16():
        xh = x_hat.clone().requires_grad_()
        out = O.d_forward(dp, xh, n, alpha, ARCH)
        g, = torch.autograd.grad(out.sum(), xh, create_graph=True)
        pen = 10.0 * torch.mean((g.norm(2, dim=(1, 2, 3)) - 1) ** 2)
        ref = dict(zip(names, torch.autograd.grad(pen, [dp[k] for k in names], allow_unused=True)))
    D.zero_grad()
    pen_f = autograd_fns.gradient_penalty(D, x_hat.to(DEV), 10.0)
    pen_f.backward()
    assert abs(pen_f.item() - pen.item()) <= 2e-3 * abs(pen.item()), (pen_f.item(), pen.item())
    params = dict(D.named_parameters())
    got = {k: params[km[k]].grad for k in names}
    report_and_check(f'gradient penalty r{res} a{alpha}', got, ref, [k for k in names if ref[k] is not None])


def test_default_epsilon_is_the_device_draw_of_the_reference():
    """D_grad_pen_loss without an injected epsilon draws torch.rand((B,1,1,1), device=images.device) after the latent
    draw and the generator pass, exactly like loss_functions.py:166-170: under a fixed CUDA seed it reproduces the
    result of injecting that draw."""
    from neuron_gan_b200.loss_functions import D_grad_pen_loss
    from neuron_gan_b200.utils import sample_latent_vec
    G, D = nets(32, 1.0)
    x = O.synthetic_images(4, 32, seed=9).to(DEV)
    gp_f = D_grad_pen_loss(G, D, Lambda=10)
    torch.manual_seed(3)
    torch.cuda.manual_seed(5)
    rng = torch.get_rng_state()
    pen_default = gp_f(x)
    torch.set_rng_state(rng)
    torch.cuda.manual_seed(5)
    eps = torch.rand((4, 1, 1, 1), device=DEV)          # the first device draw after the seed, as at :170
    pen_injected = gp_f(x, epsilon=eps)
    assert pen_default.item() == pen_injected.item()
    torch.cuda.manual_seed(6)
    torch.set_rng_state(rng)
    assert gp_f(x).item() != pen_default.item()         # and it does depend on the device generator
