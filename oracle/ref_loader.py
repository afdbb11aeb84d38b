"""Import the UNMODIFIED reference modules (test / baseline infrastructure only).

Source directory, first that exists: $NGAN_REFERENCE_DIR, /root/reference (build container), oracle/_ref (the
git-ignored copy made by oracle/make_ref.sh, which travels to the GPU box).  The reference's utils.py imports
`parse` and `matplotlib`, data/NeuronDataset.py imports `skimage` -- all absent from this image (SURVEY.md section
0 row 14); they are stubbed in sys.modules.  Nothing of the reference is modified.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = [os.environ.get('NGAN_REFERENCE_DIR'), '/root/reference', os.path.join(_HERE, '_ref')]


def ref_dir():
    for d in CANDIDATES:
        if d and os.path.isfile(os.path.join(d, 'models.py')):
            return d
    return None


def available() -> bool:
    return ref_dir() is not None


STUBS_DIR = os.path.join(os.path.dirname(_HERE), 'shim', 'stubs')


def install_stubs():
    """`parse`, `matplotlib`, `skimage`: when the real package is missing, the minimal stand-in under shim/stubs/."""
    import importlib
    for name in ('parse', 'matplotlib', 'skimage'):
        try:
            importlib.import_module(name)
        except ImportError:
            if STUBS_DIR not in sys.path:
                sys.path.append(STUBS_DIR)
            importlib.import_module(name)
    import matplotlib.pyplot  # noqa: F401


def load():
    """Returns (models, loss_functions, utils) modules of the reference."""
    d = ref_dir()
    if d is None:
        raise RuntimeError('reference not found (looked in %s)' % [c for c in CANDIDATES if c])
    sys.dont_write_bytecode = True            # the reference tree is read-only
    os.environ.setdefault('TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD', '1')
    install_stubs()
    if d not in sys.path:
        sys.path.insert(0, d)
    import models as ref_models                # noqa: E402
    import loss_functions as ref_losses        # noqa: E402
    import utils as ref_utils                  # noqa: E402
    for m in (ref_models, ref_losses, ref_utils):
        if not os.path.abspath(getattr(m, '__file__', '')).startswith(os.path.abspath(d)):
            raise RuntimeError(f'{m.__name__} was imported from {m.__file__}, not from the reference at {d}')
    return ref_models, ref_losses, ref_utils


def build_nets(res, alpha, gen_features=None, dis_features=None, image_size=512, seed=1, device='cpu'):
    """torch.manual_seed(seed) -> G -> D -> set_resolution, as train.py:114, 172, 184."""
    import torch
    ref_models, _, _ = load()
    gen_features = gen_features or [128, 64, 32, 32, 16, 16]
    dis_features = dis_features or [16, 16, 32, 32, 64, 128]
    size_init = image_size // 2 ** (len(gen_features) - 1)
    torch.manual_seed(seed)
    G = ref_models.Generator_PG(list(gen_features), image_size_init=size_init)
    D = ref_models.Discriminator_PG(list(dis_features), image_size_init=size_init)
    if res != size_init:
        G.set_resolution(res, alpha)
        D.set_resolution(res, alpha)
    return G.to(device), D.to(device)


def iteration(G, D, x, lr=1e-4, beta1=0.5, lam=10, drift=1e-3, opts=None):
    """The body of pggan_train's inner loop (train.py:356-394) with n_critic = 1, using the reference's own loss
    modules and torch.optim.Adam.  Returns (stats, d_grads, g_grads, opts)."""
    import torch
    _, ref_losses, _ = load()
    if opts is None:
        opts = (torch.optim.Adam(D.parameters(), lr=lr, betas=(beta1, 0.999)),
                torch.optim.Adam(G.parameters(), lr=lr, betas=(beta1, 0.999)))
    opt_d, opt_g = opts
    d_loss_f = ref_losses.D_W_loss(G, D, drift_epsilon=drift)
    gp_f = ref_losses.D_grad_pen_loss(G, D, Lambda=lam)
    g_loss_f = ref_losses.G_W_loss(G, D)
    D.zero_grad()
    d_loss, sr, sf = d_loss_f(x)
    pen = gp_f(x)
    d_loss += pen
    d_loss.backward()
    d_grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in D.named_parameters()}
    opt_d.step()
    G.zero_grad()
    g_loss, _ = g_loss_f(x)
    g_loss.backward()
    g_grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in G.named_parameters()}
    opt_g.step()
    stats = {'score_real': sr.item(), 'score_fake': sf.item(), 'D_loss': d_loss.item(),
             'G_loss': g_loss.item(), 'D_grad_pen': pen.item()}
    return stats, d_grads, g_grads, opts


def timed_iterations(res, alpha, batch, steps, warmup, device='cpu', keep_grads=False):
    """Time `steps` iterations of the reference's own step (after `warmup`).  Returns (images/s, ms/step)."""
    import time

    import torch
    G, D = build_nets(res, alpha, device=device)
    gen = torch.Generator().manual_seed(7)
    x = (torch.rand(batch, 1, res, res, generator=gen) * 2 - 1).to(device)
    _, ref_losses, _ = load()
    opt_d = torch.optim.Adam(D.parameters(), lr=1e-4, betas=(0.5, 0.999))
    opt_g = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.5, 0.999))
    d_loss_f = ref_losses.D_W_loss(G, D, drift_epsilon=1e-3)
    gp_f = ref_losses.D_grad_pen_loss(G, D, Lambda=10)
    g_loss_f = ref_losses.G_W_loss(G, D)

    def one():
        D.zero_grad()                                   # train.py:357-366
        d_loss, sr, sf = d_loss_f(x)
        pen = gp_f(x)
        d_loss += pen
        d_loss.backward()
        opt_d.step()
        G.zero_grad()                                   # train.py:375-385
        g_loss, _ = g_loss_f(x)
        g_loss.backward()
        opt_g.step()
        return [v.item() for v in (sr, sf, d_loss, g_loss, pen)] + [pen.item()]   # six .item() reads, train.py:389-394

    def sync():
        if str(device).startswith('cuda'):
            torch.cuda.synchronize()
    for _ in range(warmup):
        one()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    sync()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3
