"""Generator-only inference (BASELINE config 5): utils.gen_samples / gen_samples_host /
plot_gen_samples against the committed outputs of the UNMODIFIED reference's gen_samples (tests/golden/
eval_golden.pt, made by tests/golden/gen_eval_golden.py; reference utils.py:346-355, 568-610, eval.py:23-26).

Tolerance (bf16 operands, fp32 accumulation): images rel-L2 <= 2e-2 against the fp32 reference; latents bit-exact."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def eval_golden():
    return torch.load(os.path.join(HERE, 'golden', 'eval_golden.pt'), weights_only=False)


def nets(res, alpha):
    from neuron_gan_b200.train_step import build_networks
    return build_networks(res, alpha, seed=1, device=DEV)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('key', ['r512_a1.0', 'r64_a0.5'])
def test_gen_samples_matches_reference(eval_golden, key):
    from neuron_gan_b200 import utils
    ref = eval_golden[key]
    res, alpha = (512, 1.0) if key.startswith('r512') else (64, 0.5)
    G, _ = nets(res, alpha)
    utils.Latent_vecs_memo.clear()
    state = torch.get_rng_state()
    images, z = utils.gen_samples(G, N_images=ref['n'], seed=0)
    assert torch.equal(torch.get_rng_state(), state)                  # the seeded draw restores the global stream
    assert torch.equal(z.cpu(), ref['z'])                             # bit-exact latents (utils.py:57-92)
    assert (ref['n'], 512) in [k[0] for k in utils.Latent_vecs_memo]  # memoised by (size, mode, seed)
    assert tuple(images.shape) == ref['shape'] and images.dtype == torch.float32
    assert rel(images[:, 0, :16, :16].cpu(), ref['patch']) < 2e-2
    got = images.double().sum(dim=(1, 2, 3)).cpu()
    assert torch.all((got - ref['sum']).abs() < 2e-2 * ref['abssum'])
    # chunk-size invariance: every output pixel is computed the same way whatever the batch a chunk carries
    small, _ = utils.gen_samples(G, N_images=ref['n'], seed=0, chunk=3)
    assert torch.equal(small, images)
    # bf16 images (BASELINE config 5) are the fp32 images rounded
    half, _ = utils.gen_samples(G, N_images=ref['n'], seed=0, dtype=torch.bfloat16)
    assert half.dtype == torch.bfloat16 and torch.equal(half, images.to(torch.bfloat16))
    # the host path (pinned result, chunked double-buffered copies) carries the same bytes
    host = utils.gen_samples_host(G, ref['n'], seed=0, chunk=3)
    torch.cuda.synchronize()
    assert not host.is_cuda and host.is_pinned() and torch.equal(host, images.cpu())
    host16 = utils.gen_samples_host(G, ref['n'], seed=0, dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert torch.equal(host16, half.cpu())


def test_plot_gen_samples_writes_the_grid(tmp_path, eval_golden):
    """plot_gen_samples (utils.py:568-610): nearest-upsample to image_size_max, nrow = round(sqrt(N)), normalize=True,
    written as a PNG; eval.py's positional call (eval.py:26) means N_images."""
    from PIL import Image

    from neuron_gan_b200 import utils
    G, _ = nets(64, 0.5)
    path = str(tmp_path / 'grid.png')
    images = utils.plot_gen_samples(G, N_images=6, seed=0, filename=path)
    assert images.shape == (6, 1, 512, 512) and G.training                       # upsampled to image_size_max
    small = torch.nn.functional.avg_pool2d(images, 8)                             # nearest x8 then mean x8 = identity
    assert rel(small[:, 0, :16, :16], eval_golden['r64_a0.5']['patch']) < 2e-2
    im = Image.open(path)
    n_rows = 2                                                                   # round(sqrt(6)) images per row
    assert im.size == (n_rows * 512 + (n_rows + 1) * 2, 3 * 512 + 4 * 2)         # torchvision grid, padding 2
    path2 = str(tmp_path / 'grid2.png')
    images2 = utils.plot_gen_samples(G, 6, seed=0, filename=path2)                # eval.py:26's calling convention
    assert torch.equal(images2, images) and os.path.getsize(path2) == os.path.getsize(path)


def test_from_state_dict_puts_the_network_on_the_gpu(tmp_path):
    """eval.py:23: Generator_PG.from_state_dict(path) with no device -> a network that can run."""
    from neuron_gan_b200 import models, utils
    G, D = nets(32, 1.0)
    path = str(tmp_path / 'w.pth')
    utils.Checkpointer(G, D, 1e-4, path, N_epochs=4, verbose=False).save_state(2)
    G2 = models.Generator_PG.from_state_dict(path, verbose=False)
    assert next(G2.parameters()).is_cuda and G2.alpha.is_cuda and G2.image_size == 32
    a, _ = utils.gen_samples(G, 5, seed=3)
    b, _ = utils.gen_samples(G2, 5, seed=3)
    assert torch.equal(a, b)
    G3 = models.Generator_PG.from_state_dict(path, device=torch.device('cuda'), verbose=False)
    assert torch.equal(utils.gen_samples(G3, 5, seed=3)[0], a)


def test_weights_written_through_data_need_invalidate():
    """engine.invalidate: in-place writes through `.data` do not bump the parameter version; after
    invalidate() the kernels see the new weights, in eager calls and in TrainStep's replay decision."""
    from neuron_gan_b200 import engine, utils
    G, _ = nets(32, 1.0)
    a, _ = utils.gen_samples(G, 2, seed=1)
    w = G.layers[4].weight
    v0 = w._version
    w.data.mul_(0.5)
    assert w._version == v0                          # this is why the hook is needed
    engine.invalidate(G)
    b, _ = utils.gen_samples(G, 2, seed=1)
    assert not torch.equal(a, b)
    with torch.no_grad():
        w.mul_(2.0)                                  # a write through the parameter itself needs nothing
    c, _ = utils.gen_samples(G, 2, seed=1)
    assert torch.equal(a, c)
