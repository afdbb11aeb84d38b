"""The oracle (oracle/pggan_oracle.py) against the committed outputs of the reference itself
(tests/golden/pggan_step_golden.pt, made by tests/golden/gen_golden.py). CPU only."""
import pytest
import torch

from oracle import pggan_oracle as O

ARCH = O.Arch()
CASES = ['r16_a1.0_b16', 'r32_a0.5_b4', 'r32_a1.0_b4', 'r64_a0.5_b64', 'r64_a1.0_b4',
         'r128_a0.25_b2', 'r128_a1.0_b2', 'r256_a1.0_b1', 'r512_a0.5_b1', 'r512_a1.0_b2',
         'r16_a1.0_b1', 'r64_a1.0_b3', 'r128_a0.25_b5', 'r256_a0.5_b3']        # ragged last batches


def close(a, b, rel=2e-5, abs_=1e-7):
    return abs(a - b) <= rel * max(abs(a), abs(b)) + abs_


def test_init_params_match_reference(golden):
    gp, dp = O.build_params(ARCH, seed=1)
    for name, ref in golden['init']['g'].items():
        assert tuple(gp[name].shape) == ref['shape']
        assert torch.equal(gp[name].flatten()[:8], ref['head']), name
        assert close(gp[name].double().sum().item(), ref['sum'], 1e-9), name
    for name, ref in golden['init']['d'].items():
        assert torch.equal(dp[name].flatten()[:8], ref['head']), name
        assert close(dp[name].double().abs().sum().item(), ref['abssum'], 1e-9), name
    n_g = sum(v.numel() for v in gp.values())
    n_d = sum(v.numel() for v in dp.values())
    assert (n_g, n_d) == (17093152, 494273)


def test_scales_and_gain():
    assert abs(O.he_gain(0.2) - 1.3867504905630728) < 1e-15
    w = torch.zeros(16, 128, 3, 3)
    assert abs(O.conv_scale(w, 0.2) - 1.3867504905630728 / (128 * 9) ** 0.5) < 1e-15


def test_bilinear_taps_are_the_fixed_pattern():
    # x2 bilinear = weights .25/.75 with clamped indices; x0.5 == 2x2 mean
    x = torch.arange(16.).reshape(1, 1, 4, 4)
    up = O.up2(x)
    row = x[0, 0, 0]
    exp = torch.stack([row[0], .75 * row[0] + .25 * row[1], .25 * row[0] + .75 * row[1],
                       .75 * row[1] + .25 * row[2]])
    col0 = torch.tensor([.75, .25])  # rows 0 (clamped: all weight on row 0) handled below
    assert torch.allclose(up[0, 0, 0, :4], exp)
    imp = torch.zeros(1, 1, 8, 8)
    imp[0, 0, 4, 4] = 1
    vals = sorted(set(round(v, 6) for v in O.up2(imp).flatten().tolist() if v > 0))
    assert vals == [0.0625, 0.1875, 0.5625]
    y = torch.randn(2, 3, 8, 8)
    assert torch.equal(O.down2_bilinear(y), torch.nn.functional.avg_pool2d(y, 2))


def test_pixelnorm_double_backward_formulas():
    """hand-derived PixelNorm backward / double-backward vs autograd (fp64)."""
    torch.manual_seed(0)
    x = torch.randn(2, 16, 3, 3, dtype=torch.float64, requires_grad=True)
    g = torch.randn_like(x, requires_grad=True)
    gh = torch.randn_like(x)
    y = O.pixel_norm(x)
    r = 1 / torch.sqrt((x ** 2).mean(1, keepdim=True) + 1e-8)
    dx, = torch.autograd.grad(y, x, g, create_graph=True)
    t = (g * y).mean(1, keepdim=True)
    assert torch.allclose(dx, r * (g - y * t), atol=1e-12)
    cg, cx = torch.autograd.grad(dx, (g, x), gh)
    u = (gh * y).mean(1, keepdim=True)
    v = (gh * g).mean(1, keepdim=True)
    assert torch.allclose(cg, r * (gh - y * u), atol=1e-12)
    assert torch.allclose(cx, -r * r * (t * gh + u * g + (v - 3 * u * t) * y), atol=1e-12)


@pytest.mark.parametrize('key', CASES)
def test_iteration_matches_reference(golden, key):
    ref = golden['cases'][key]
    res, alpha, batch = ref['res'], ref['alpha'], ref['batch']
    n = O.n_layers_for(res, ARCH)
    tr = O.Trainer(ARCH, seed=1, res=res, alpha=alpha)
    # state-dict key translation covers exactly the reference's keys
    assert sorted(O.g_key_map(n, alpha < 1, ARCH).values()) == ref['g_keys']
    assert sorted(list(O.d_key_map(n, alpha < 1, ARCH).values()) + ['alpha']) == ref['d_keys']
    x = O.synthetic_images(batch, res)
    assert close(x.double().sum().item(), ref['x']['sum'], 1e-12)
    rng = torch.get_rng_state()
    z1, z2, eps, z3 = tr.draw(batch)
    assert torch.equal(eps.flatten(), ref['draws']['eps'])               # bit-exact draws
    assert torch.equal(z1[:2], ref['draws']['z1_rows']) and torch.equal(z3[:2], ref['draws']['z3_rows'])
    assert close(z2.double().sum().item(), ref['draws']['z2']['sum'], 1e-9)
    with torch.no_grad():
        img = O.g_forward(tr.gp, z1, n, alpha, ARCH)
        assert torch.allclose(img[:2, 0, :8, :8], ref['g_img_patch'], rtol=1e-4, atol=1e-6)
        assert torch.allclose(O.d_forward(tr.dp, x, n, alpha, ARCH).flatten(), ref['d_real'], rtol=1e-4, atol=1e-6)
    pen, g1 = O.grad_penalty(tr.gp, tr.dp, x, z2, eps, n, alpha, ARCH, return_grad=True)
    assert torch.allclose(g1[:2, 0, :8, :8], ref['gp_grad_patch'], rtol=1e-3, atol=1e-8)
    assert torch.allclose(g1.norm(2, dim=(1, 2, 3)), ref['gp_grad_norms'], rtol=1e-4)
    torch.set_rng_state(rng)
    stats = tr.iteration(x)
    for k, v in ref['stats'].items():
        assert close(stats[k], v, 2e-5, 2e-7), (k, stats[k], v)
    for grads, refg, inactive in ((tr.last_d_grads, ref['d_grads'], ref['d_inactive']),
                                  (tr.last_g_grads, ref['g_grads'], ref['g_inactive'])):
        assert sorted(k for k, g in grads.items() if g is not None) == sorted(refg.keys())
        assert not (set(grads.keys()) & set(inactive))
        for k, s in refg.items():
            assert close(grads[k].double().norm().item(), s['norm'], 1e-3, 1e-9), k
            assert torch.allclose(grads[k].flatten()[:8], s['head'], rtol=5e-3, atol=1e-7), k
    # parameters after the two Adam steps (isolated |g|~1e-8 elements may move differently)
    for params, refp in ((tr.gp, ref['g_after']), (tr.dp, ref['d_after'])):
        for k, s in refp.items():
            p = params[k].detach()
            assert abs(p.double().sum().item() - s['sum']) <= 2.5e-4 * max(1, p.numel() ** 0.5), k
            assert torch.allclose(p.flatten()[:8], s['head'], rtol=0, atol=2.1e-4), k


def test_adam_matches_torch_optim():
    torch.manual_seed(3)
    p0 = {'a': torch.randn(7, 5), 'b': torch.randn(11)}
    mine = {k: v.clone() for k, v in p0.items()}
    theirs = [torch.nn.Parameter(v.clone()) for v in p0.values()]
    opt = torch.optim.Adam(theirs, lr=1e-3, betas=(0.5, 0.999))
    adam = O.Adam(mine, lr=1e-3, beta1=0.5)
    for step in range(5):
        grads = {k: torch.randn_like(v) for k, v in p0.items()}
        if step == 2:
            grads['b'] = None                     # skipped parameter: its step count must not advance
        for p, g in zip(theirs, grads.values()):
            p.grad = g
        opt.step()
        adam.step(grads)
    for p, k in zip(theirs, mine):
        assert torch.allclose(p.data, mine[k], rtol=1e-6, atol=1e-7), k


@pytest.mark.parametrize('key', ['r32_a0.5_b4_n2', 'r32_a0.5_b4_n0', 'r64_a1.0_b3_n3', 'r32_a0.5_b4_n1_lam0'])
def test_critic_loop_variants(key):
    """The oracle's n_critic = 2 / 3 / 0 and grad_pen_lambda = 0 iterations against the unmodified reference
    (tests/golden/gen_ncritic_golden.py): statistics, stream position, parameters after the iteration."""
    import os
    ref = torch.load(os.path.join(os.path.dirname(__file__), 'golden', 'ncritic_golden.pt'), weights_only=False)['cases'][key]
    tr = O.Trainer(ARCH, seed=1, res=ref['res'], alpha=ref['alpha'], lam=ref['lam'])
    x = O.synthetic_images(ref['batch'], ref['res'], seed=ref['image_seed'])
    torch.manual_seed(ref['draw_seed'])
    stats = tr.iteration(x, n_critic=ref['n_critic'])
    assert torch.rand(1).item() == ref['next_draw']
    for k, v in ref['stats'].items():
        assert close(stats[k], v, 2e-5, 2e-7), (k, stats[k], v)
    for params, refp in ((tr.gp, ref['g_after']), (tr.dp, ref['d_after'])):
        for k, s in refp.items():
            p = params[k].detach()
            assert abs(p.double().sum().item() - s['sum']) <= 2.5e-4 * max(1, ref['n_critic']) * max(1, p.numel() ** 0.5), k
            assert torch.allclose(p.flatten()[:8], s['head'], rtol=0, atol=2.1e-4 * max(1, ref['n_critic'])), k
