"""Import the UNMODIFIED reference (/root/reference) in the build container.

Only used by tests/golden/gen_golden.py and tests/test_oracle_vs_reference.py.  The GPU box has no
/root/reference: everything that runs there uses the committed fixtures in tests/golden/ instead.

The reference's utils.py imports `parse` and `matplotlib` (absent from this image, SURVEY.md
section 8c); they are stubbed in sys.modules. Nothing in /root/reference is modified or copied.
"""
import os
import sys
import types

REF_DIR = os.environ.get('NGAN_REFERENCE_DIR', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, 'models.py'))


def load():
    """Returns (models, loss_functions, utils) modules of the reference."""
    if not available():
        raise RuntimeError(f'reference not found under {REF_DIR}')
    sys.dont_write_bytecode = True            # the reference tree is read-only
    os.environ.setdefault('TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD', '1')
    if 'parse' not in sys.modules:
        m = types.ModuleType('parse')
        m.parse = lambda *a, **k: None
        sys.modules['parse'] = m
    if 'matplotlib' not in sys.modules:
        mpl = types.ModuleType('matplotlib')
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType('matplotlib.pyplot')
        mpl.pyplot = plt
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import models as ref_models                # noqa: E402
    import loss_functions as ref_losses        # noqa: E402
    import utils as ref_utils                  # noqa: E402
    return ref_models, ref_losses, ref_utils


def build_nets(res, alpha, gen_features=None, dis_features=None, image_size=512, seed=1):
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
    return G, D


def iteration(G, D, x, lr=1e-4, beta1=0.5, lam=10, drift=1e-3, opts=None):
    """The body of pggan_train's inner loop (train.py:356-394) with n_critic = 1, using the
    reference's own loss modules and torch.optim.Adam."""
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
