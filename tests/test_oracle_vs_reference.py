"""Live comparison oracle <-> UNMODIFIED reference; runs only where /root/reference exists (the
build container). On the GPU box these are skipped and the committed fixtures take over."""
import pytest
import torch

import ref_harness as rh
from oracle import pggan_oracle as O

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference tree not present')
ARCH = O.Arch()


@pytest.mark.parametrize('res,alpha,batch', [(16, 1.0, 3), (32, 0.3, 2), (64, 1.0, 2), (128, 0.7, 1)])
def test_forward_and_losses(res, alpha, batch):
    G, D = rh.build_nets(res, alpha, seed=5)
    n = O.n_layers_for(res, ARCH)
    gp = O.g_state_to_named(G.state_dict(), n, alpha < 1, ARCH)
    dp = O.d_state_to_named(D.state_dict(), n, alpha < 1, ARCH)
    x = O.synthetic_images(batch, res, seed=11)
    z = O.sample_latent((batch, 512))
    with torch.no_grad():
        assert torch.allclose(G(z), O.g_forward(gp, z, n, alpha, ARCH), rtol=1e-5, atol=1e-6)
        assert torch.allclose(D(x), O.d_forward(dp, x, n, alpha, ARCH), rtol=1e-5, atol=1e-6)


def test_latent_sampler_bit_exact():
    _, _, ref_utils = rh.load()
    torch.manual_seed(123)
    a = ref_utils.sample_latent_vec((5, 512))
    torch.manual_seed(123)
    b = O.sample_latent((5, 512))
    assert torch.equal(a, b)


def test_small_architecture_iteration():
    """A non-default architecture (3 levels, 4x4 start) exercises the generic key maps."""
    gen_f, dis_f = [32, 16, 8], [8, 16, 32]
    arch = O.Arch(gen_features=gen_f, dis_features=dis_f, image_size=16)
    G, D = rh.build_nets(16, 0.5, gen_f, dis_f, image_size=16, seed=2)
    rng = torch.get_rng_state()
    x = O.synthetic_images(4, 16)
    stats, d_grads, g_grads, _ = rh.iteration(G, D, x)
    tr = O.Trainer(arch, seed=2, res=16, alpha=0.5)
    torch.set_rng_state(rng)
    mine = tr.iteration(x)
    for k in stats:
        assert abs(stats[k] - mine[k]) <= 2e-5 * max(1, abs(stats[k])), k
    dkm = O.d_key_map(3, True, arch)
    for name, key in dkm.items():
        if d_grads[key] is not None:
            assert torch.allclose(d_grads[key], tr.last_d_grads[name], rtol=1e-3, atol=1e-7), name
