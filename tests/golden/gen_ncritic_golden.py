"""Generate tests/golden/ncritic_golden.pt by running the UNMODIFIED reference on CPU: the inner iteration of
pggan_train (train.py:355-394) for N_D_steps = 2 and N_D_steps = 0 (the adapt_critic case, train.py:336-340, where
the losses are only evaluated for monitoring, train.py:370-374).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_ncritic_golden.py

Recorded per case: the five statistics train.py accumulates (of the LAST critic round, train.py:390) and checksums /
the first eight elements of every parameter of both networks after the iteration."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import ref_harness as rh                      # noqa: E402
from oracle import pggan_oracle as O          # noqa: E402

CASES = [(32, 0.5, 4, 2, 10), (32, 0.5, 4, 0, 10), (64, 1.0, 3, 3, 10),     # (res, alpha, batch, N_D_steps, Lambda)
         (32, 0.5, 4, 1, 0)]       # grad_pen_lambda = 0: D_grad_pen_loss returns 0 without drawing (loss_functions.py:159)


def summarize(t):
    f = t.detach().flatten().double()
    return {'norm': f.norm().item(), 'sum': f.sum().item(), 'abssum': f.abs().sum().item(),
            'head': t.detach().flatten()[:8].clone()}


def run_case(res, alpha, batch, n_d, lam):
    _, ref_losses, _ = rh.load()
    arch = O.Arch()
    n = O.n_layers_for(res, arch)
    G, D = rh.build_nets(res, alpha)
    gkm, dkm = O.g_key_map(n, alpha < 1, arch), O.d_key_map(n, alpha < 1, arch)
    x = O.synthetic_images(batch, res, seed=61)
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_g = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    d_loss_f = ref_losses.D_W_loss(G, D, drift_epsilon=1e-3)
    gp_f = ref_losses.D_grad_pen_loss(G, D, Lambda=lam)
    g_loss_f = ref_losses.G_W_loss(G, D)
    torch.manual_seed(123)                      # the draws of the iteration (the test re-seeds the same way)
    for _ in range(n_d):                        # train.py:356-366
        D.zero_grad()
        d_loss, sr, sf = d_loss_f(x)
        pen = gp_f(x)
        d_loss += pen
        d_loss.backward()
        opt_d.step()
    if n_d == 0:                                # train.py:370-374
        d_loss, sr, sf = d_loss_f(x)
        pen = gp_f(x)
        d_loss += pen
    G.zero_grad()                               # train.py:376-386
    g_loss, _ = g_loss_f(x)
    g_loss.backward()
    opt_g.step()
    gs, ds = G.state_dict(), D.state_dict()
    next_draw = torch.rand(1).item()            # where the CPU stream stands after the iteration
    return {'res': res, 'alpha': alpha, 'batch': batch, 'n_critic': n_d, 'lam': lam, 'draw_seed': 123, 'image_seed': 61,
            'next_draw': next_draw,
            'stats': {'score_real': sr.item(), 'score_fake': sf.item(), 'D_loss': d_loss.item(),
                      'G_loss': g_loss.item(), 'D_grad_pen': float(pen)},
            'g_after': {k: summarize(gs[v]) for k, v in gkm.items()},
            'd_after': {k: summarize(ds[v]) for k, v in dkm.items()}}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {'torch': torch.__version__, 'cases': {}}
    for res, alpha, batch, n_d, lam in CASES:
        key = f'r{res}_a{alpha}_b{batch}_n{n_d}' + ('' if lam == 10 else f'_lam{lam}')
        out['cases'][key] = run_case(res, alpha, batch, n_d, lam)
        print(key, out['cases'][key]['stats'])
    path = os.path.join(HERE, 'ncritic_golden.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
