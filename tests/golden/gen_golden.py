"""Generate tests/golden/pggan_step_golden.pt by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_golden.py [--add]

(--add keeps the cases already in the fixture byte for byte and only generates the missing ones.)

For every (resolution, alpha, batch) case it builds the reference Generator_PG / Discriminator_PG
after torch.manual_seed(1) (order G, D: train.py:114, 172, 184), calls set_resolution, and runs the
body of pggan_train's inner iteration (train.py:356-394) with the reference's own loss modules and
torch.optim.Adam on synthetic U[-1,1) images.  It records the RNG draws (z, z, eps, z), the five
loss statistics, slices / norms of network outputs and parameter gradients, and parameter
checksums after the Adam updates.  The fixture is what pins oracle/pggan_oracle.py (CPU tests) and
what the CUDA path is compared with on the GPU box, where /root/reference does not exist.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import ref_harness as rh                      # noqa: E402
from oracle import pggan_oracle as O          # noqa: E402

CASES = [  # (res, alpha, batch)
    (16, 1.0, 16),     # BASELINE config 1
    (32, 0.5, 4),
    (32, 1.0, 4),
    (64, 0.5, 64),     # BASELINE config 2
    (64, 1.0, 4),
    (128, 0.25, 2),
    (128, 1.0, 2),
    (256, 1.0, 1),
    (512, 0.5, 1),
    (512, 1.0, 2),     # BASELINE config 3 shape (reduced batch)
    (512, 1.0, 16),    # BASELINE config 3 at its per-GPU batch (16 images per GPU)
    # ragged batches: what the reference's DataLoader delivers as the last batch of an epoch (drop_last=False)
    (16, 1.0, 1),
    (64, 1.0, 3),
    (128, 0.25, 5),
    (256, 0.5, 3),
]


def summarize(t: torch.Tensor):
    f = t.detach().flatten().double()
    return {'norm': f.norm().item(), 'sum': f.sum().item(), 'abssum': f.abs().sum().item(),
            'head': t.detach().flatten()[:8].clone(), 'shape': tuple(t.shape)}


def run_case(res, alpha, batch):
    _, ref_losses, ref_utils = rh.load()
    arch = O.Arch()
    n = O.n_layers_for(res, arch)
    G, D = rh.build_nets(res, alpha)
    gkm, dkm = O.g_key_map(n, alpha < 1, arch), O.d_key_map(n, alpha < 1, arch)
    inv_g, inv_d = {v: k for k, v in gkm.items()}, {v: k for k, v in dkm.items()}
    x = O.synthetic_images(batch, res)
    out = {'res': res, 'alpha': alpha, 'batch': batch, 'seed': 1, 'x': summarize(x)}

    # replay the draws the iteration will make, to record them (utils.py:57-92, loss_functions.py:170)
    rng = torch.get_rng_state()
    z1 = ref_utils.sample_latent_vec((batch, 512))
    z2 = ref_utils.sample_latent_vec((batch, 512))
    eps = torch.rand((batch, 1, 1, 1))
    z3 = ref_utils.sample_latent_vec((batch, 512))
    out['draws'] = {'z1': summarize(z1), 'z2': summarize(z2), 'z3': summarize(z3), 'eps': eps.flatten().clone(),
                    'z1_rows': z1[:2].clone(), 'z3_rows': z3[:2].clone()}

    # forward-only observables
    with torch.no_grad():
        img = G(z1)
        out['g_img'] = summarize(img)
        out['g_img_patch'] = img[:2, 0, :8, :8].clone()
        out['d_real'] = D(x).flatten().clone()
        out['d_fake'] = D(img).flatten().clone()
        x_tilde = G(z2)
    x_hat = (eps * x + (1 - eps) * x_tilde).requires_grad_()
    g1 = torch.autograd.grad(D(x_hat).sum(), x_hat)[0]
    out['gp_grad'] = summarize(g1)
    out['gp_grad_patch'] = g1[:2, 0, :8, :8].clone()
    out['gp_grad_norms'] = g1.norm(2, dim=(1, 2, 3)).clone()

    torch.set_rng_state(rng)
    stats, d_grads, g_grads, _ = rh.iteration(G, D, x)
    out['stats'] = stats
    out['d_grads'] = {inv_d[k]: summarize(v) for k, v in d_grads.items() if v is not None}
    out['g_grads'] = {inv_g[k]: summarize(v) for k, v in g_grads.items() if v is not None}
    out['d_inactive'] = sorted(inv_d[k] for k, v in d_grads.items() if v is None)
    out['g_inactive'] = sorted(inv_g[k] for k, v in g_grads.items() if v is None)
    gs, ds = G.state_dict(), D.state_dict()
    out['g_after'] = {k: summarize(gs[v]) for k, v in gkm.items()}
    out['d_after'] = {k: summarize(ds[v]) for k, v in dkm.items()}
    out['g_keys'] = sorted(gs.keys())
    out['d_keys'] = sorted(ds.keys())
    return out


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    path = os.path.join(HERE, 'pggan_step_golden.pt')
    golden = {'torch': torch.__version__, 'cases': {}}
    if '--add' in sys.argv and os.path.exists(path):
        golden['cases'] = torch.load(path, weights_only=False)['cases']
    # initial parameters (seed 1) are pinned by checksum so that the GPU box can rebuild them
    gp, dp = O.build_params(O.Arch(), seed=1)
    G, D = rh.build_nets(16, 1.0)
    gs, ds = G.state_dict(), D.state_dict()
    for k, v in O.g_key_map(1, False, O.Arch()).items():
        assert torch.equal(gp[k], gs[v]), k
    for k, v in O.d_key_map(1, False, O.Arch()).items():
        assert torch.equal(dp[k], ds[v]), k
    golden['init'] = {'g': {k: summarize(v) for k, v in gp.items()},
                      'd': {k: summarize(v) for k, v in dp.items()}}
    for res, alpha, batch in CASES:
        key = f'r{res}_a{alpha}_b{batch}'
        if key in golden['cases']:
            continue
        golden['cases'][key] = run_case(res, alpha, batch)
        print(key, golden['cases'][key]['stats'])
    torch.save(golden, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
