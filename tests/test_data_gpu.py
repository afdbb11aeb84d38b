"""On-device image pipeline (csrc/augment.cu through the C ABI) against the committed outputs of the unmodified
reference DatasetIterator and against the oracle at the full 768 -> 512 geometry."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as do

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'data_pipeline_golden.pt')


def _close(got, want, crop):
    """Nearest-neighbour resampling is discontinuous (oracle/data_oracle.py header): a source coordinate on a
    half-integer to within float rounding may pick the other neighbour.  Allow 1 source pixel in 4000 per image to do
    so, each reaching at most 2x2 antialiased outputs; everything else agrees to float rounding of the filter sums."""
    d = np.abs(got - want)
    allowed = got.shape[0] * 4 * int(np.ceil(2.5e-4 * crop * crop))
    assert (d > 1e-5).sum() <= allowed and np.median(d) <= 2e-6, (crop, d.max(), (d > 1e-5).sum(), allowed)


def test_reference_batches():
    from neuron_gan_b200 import data
    g = torch.load(GOLDEN, weights_only=True)
    ds = data.NeuronImages(g['canvases'], g['image_size_max'], True, g['translate'])
    it = data.DatasetIterator(ds, g['batch_size'], 'cuda')
    for ep in g['epochs']:
        ds.set_image_size(ep['size'])
        torch.manual_seed(ep['seed'])
        got = [b.cpu().clone() for b in it]
        assert [tuple(b.shape) for b in got] == [tuple(b.shape) for b in ep['batches']]
        for a, b in zip(got, ep['batches']):
            _close(a.numpy(), b.numpy(), g['image_size_max'])
    ds2 = data.NeuronImages(g['canvases'], g['image_size_max'], False)
    ds2.set_image_size(16)
    for a, b in zip(data.DatasetIterator(ds2, g['batch_size'], 'cuda'), g['plain16']):
        assert (a.cpu() - b).abs().max().item() <= 1e-6


@pytest.mark.parametrize('size', [512, 256, 64, 16])
def test_full_geometry_vs_oracle(size):
    """BASELINE geometry: 512 px images on 768 px canvases, every phase's output size."""
    from neuron_gan_b200 import data
    rng = np.random.RandomState(size)
    yy, xx = np.mgrid[0:768, 0:768].astype(np.float32)
    cv = np.stack([np.clip(0.5 + 0.4 * np.sin(xx / (7 + k)) * np.cos(yy / (5 + k)) + 0.1 * rng.randn(768, 768), 0, 1)
                   for k in range(3)]).astype(np.float32)
    ds = data.NeuronImages(torch.from_numpy(cv), 512, True, 0.05)
    ds.set_image_size(size)
    torch.manual_seed(100 + size)
    got = [b.cpu().numpy().copy() for b in data.DatasetIterator(ds, 2, 'cuda')]
    torch.manual_seed(100 + size)
    want = [b for b, _ in do.epoch_batches(cv, 2, 512, size, 0.05)]
    assert [b.shape for b in got] == [b.shape for b in want] == [(2, 1, size, size), (1, 1, size, size)]
    for a, b in zip(got, want):
        _close(a, b, 512)
        assert a.min() >= -1.0 and a.max() <= 1.0


@pytest.mark.parametrize('canvas,crop,size', [(97, 64, 64), (97, 64, 16), (75, 50, 50), (75, 50, 20), (100, 64, 32)])
def test_odd_geometries_vs_oracle(canvas, crop, size):
    """canvas - crop odd: the crop window and its mirror under the vertical flip differ by one row (the staging window
    of csrc/augment.cu then holds crop + 1 rows); sizes that are not powers of two."""
    from neuron_gan_b200 import data
    cv = torch.rand(5, canvas, canvas, generator=torch.Generator().manual_seed(canvas + size)).numpy()
    ds = data.NeuronImages(torch.from_numpy(cv), crop, True, 0.05)
    ds.set_image_size(size)
    torch.manual_seed(canvas * size)
    got = [b.cpu().numpy().copy() for b in data.DatasetIterator(ds, 3, 'cuda')]
    torch.manual_seed(canvas * size)
    want = [b for b, _ in do.epoch_batches(cv, 3, crop, size, 0.05)]
    for a, b in zip(got, want):
        assert a.shape == b.shape
        _close(a, b, crop)


def test_rank_sharding_matches_single_process():
    from neuron_gan_b200 import data
    cv = torch.rand(8, 96, 96, generator=torch.Generator().manual_seed(0))
    ds = data.NeuronImages(cv, 64, True, 0.05)
    ds.set_image_size(32)
    torch.manual_seed(5)
    whole = [b.cpu().clone() for b in data.DatasetIterator(ds, 4, 'cuda')]
    parts = []
    for r in range(2):
        torch.manual_seed(5)
        parts.append([b.cpu().clone() for b in data.DatasetIterator(ds, 4, 'cuda', rank=r, world=2)])
    for w, p0, p1 in zip(whole, parts[0], parts[1]):
        assert torch.equal(torch.cat([p0, p1]), w)


def test_feeds_the_training_step():
    """A batch from the iterator goes straight into TrainStep (device tensor, no host round trip)."""
    from neuron_gan_b200 import data
    from neuron_gan_b200.train_step import TrainStep, build_networks
    G, D = build_networks(32, 0.5, seed=1, device='cuda')
    ds = data.NeuronImages(torch.rand(8, 768, 768, generator=torch.Generator().manual_seed(2)), 512, True, 0.05)
    ds.set_image_size(32)
    step = TrainStep(G, D)
    n = 0
    for x in data.DatasetIterator(ds, 4, 'cuda'):
        assert x.shape == (4, 1, 32, 32)
        stats = step(x)
        assert torch.isfinite(stats).all()
        n += 1
    assert n == 2
