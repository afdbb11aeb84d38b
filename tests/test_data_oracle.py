"""CPU checks of the image-pipeline oracle (oracle/data_oracle.py) and of the product's host-side logic
(neuron_gan_b200/data.py: parameter draws, antialias taps, iterator bookkeeping)."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as do

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'data_pipeline_golden.pt')


@pytest.fixture(scope='module')
def golden():
    return torch.load(GOLDEN, weights_only=True)


def test_oracle_reproduces_reference_batches(golden):
    """Outputs of the unmodified reference NeuronDataset + DatasetIterator (tests/golden/gen_data_golden.py).
    Nearest resampling: allow at most 1 pixel in 1000 beyond 1e-5 (none observed), everything within 1e-6 else."""
    cv = golden['canvases'].numpy()
    for ep in golden['epochs']:
        torch.manual_seed(ep['seed'])
        got = list(do.epoch_batches(cv, golden['batch_size'], golden['image_size_max'], ep['size'],
                                    golden['translate']))
        assert len(got) == len(ep['batches']) == 3
        for (g, _), want in zip(got, ep['batches']):
            assert g.shape == tuple(want.shape)                    # ragged last batch included
            d = np.abs(g - want.numpy())
            assert (d > 1e-5).mean() <= 1e-3 and np.median(d) <= 1e-6, (ep['size'], d.max())


def test_oracle_without_augmentations(golden):
    cv = golden['canvases'].numpy()
    for (g, _), want in zip(do.epoch_batches(cv, 2, 64, 16, 0.05, augmentations=False), golden['plain16']):
        assert np.abs(g - want.numpy()).max() <= 1e-6


def test_host_param_draw_follows_the_reference_rng_order():
    """Same generator consumption and same values as the oracle's restatement of torchvision's get_params."""
    from neuron_gan_b200 import data
    for seed in range(8):
        torch.manual_seed(seed)
        row = data.draw_augment_params(96, 0.05)
        s1 = torch.get_rng_state()
        torch.manual_seed(seed)
        p = do.draw_params(96, 0.05)
        assert torch.equal(s1, torch.get_rng_state())
        th = np.asarray(do.inverse_affine_matrix(p['angle'], p['tx'], p['ty']), dtype=np.float32) / np.float32(48)
        assert np.array_equal(row[:6], th)
        assert row[6] == float(p['flip']) and row[7] == np.float32(p['b']) and row[8] == np.float32(p['c'])
        assert row[9] == np.float32(1.0 - p['c']) and row[10] == p['order'] and row[11] == 0


def test_host_param_draw_matches_torchvision_live():
    tv = pytest.importorskip('torchvision')
    from neuron_gan_b200 import data
    T = tv.transforms
    for seed in range(4):
        torch.manual_seed(seed)
        angle, (tx, ty), _, _ = T.RandomAffine.get_params([-180.0, 180.0], (0.05, 0.05), None, None, [96, 96])
        flip = bool(torch.rand(1) < 0.5)
        fn_idx, b, c, _, _ = T.ColorJitter.get_params((0.75, 1.25), (0.75, 1.25), None, None)
        torch.manual_seed(seed)
        row = data.draw_augment_params(96, 0.05)
        m = tv.transforms.functional._get_inverse_affine_matrix([0.0, 0.0], angle, [float(tx), float(ty)], 1.0,
                                                                 [0.0, 0.0])
        assert np.array_equal(row[:6], np.asarray(m, dtype=np.float32) / np.float32(48))
        assert row[6] == float(flip) and row[7] == np.float32(b) and row[8] == np.float32(c)
        order = 0 if fn_idx.tolist().index(0) < fn_idx.tolist().index(1) else 1
        assert row[10] == order


@pytest.mark.parametrize('sizes', [(64, 32), (64, 8), (512, 16), (512, 256), (60, 17), (100, 33), (64, 64)])
def test_antialias_taps(sizes):
    """Product taps == oracle taps, and both reproduce ATen's interpolate(antialias=True) on a random image."""
    from neuron_gan_b200 import data
    n_in, n_out = sizes
    first, count, weight = data.aa_taps(n_in, n_out)
    for i, (lo, w) in enumerate(do.aa_weights(n_in, n_out)):
        dense = np.zeros(n_in, dtype=np.float32)            # the product drops taps of weight exactly 0
        dense[first[i]:first[i] + count[i]] = weight[i, :count[i]]
        want = np.zeros(n_in, dtype=np.float32)
        want[lo:lo + len(w)] = w
        assert np.abs(dense - want).max() < 1e-6
        assert weight[i, count[i]:].sum() == 0
    if n_in == n_out:
        assert (count == 1).all() and (first == np.arange(n_in)).all() and (weight == 1).all()
    x = torch.rand(1, 1, n_in, n_in, generator=torch.Generator().manual_seed(n_in + n_out))
    want = torch.nn.functional.interpolate(x, size=[n_out, n_out], mode='bilinear', align_corners=False,
                                           antialias=True)[0, 0].numpy()
    got = np.zeros((n_out, n_out), dtype=np.float64)
    img = x[0, 0].numpy().astype(np.float64)
    for oy in range(n_out):
        for ox in range(n_out):
            wy = weight[oy, :count[oy]].astype(np.float64)
            wx = weight[ox, :count[ox]].astype(np.float64)
            got[oy, ox] = wy @ img[first[oy]:first[oy] + count[oy], first[ox]:first[ox] + count[ox]] @ wx
    assert np.abs(got - want).max() < 2e-6
    if n_out < n_in:
        assert np.abs(do.resize_aa(x[0, 0].numpy(), n_out) - want).max() < 2e-6


def test_dataset_mirror_bookkeeping():
    from neuron_gan_b200 import data
    from neuron_gan_b200._lib import NganError
    ds = data.NeuronImages(torch.rand(3, 1, 96, 96), image_size=64, im_translation=0.05)
    assert len(ds) == 3 and ds.image_size_max == 64 and ds.canvases.shape == (3, 96, 96)
    ds.set_image_size(16)
    assert ds.image_size == 16
    with pytest.raises(AssertionError):
        ds.set_image_size(128)                      # data/NeuronDataset.py:151
    with pytest.raises(ValueError):
        data.NeuronImages(torch.rand(3, 3, 96, 96), image_size=64)
    with pytest.raises(NganError):                  # no CPU fallback
        data.DatasetIterator(ds, 2, torch.device('cpu'))
    # from a loaded reference dataset: a list of [1, P, P] tensors ([1, 1, P, P] once its own iterator has run)
    import types
    ref_like = types.SimpleNamespace(images=[torch.rand(1, 96, 96), torch.rand(1, 1, 96, 96)], image_size_max=64,
                                     load_all=True)
    ds2 = data.NeuronImages.from_dataset(ref_like, im_translation=0.05)
    assert ds2.canvases.shape == (2, 96, 96) and ds2.image_size == 64 and ds2.im_translation == 0.05
    assert torch.equal(ds2.canvases[1], ref_like.images[1][0, 0])
    with pytest.raises(Exception, match='only possible when all images are loaded'):
        data.NeuronImages.from_dataset(types.SimpleNamespace(images=[], image_size_max=64, load_all=False))


@pytest.mark.parametrize('seed,warm,n', [(0, 0, 1), (1, 5, 16), (2, 617, 64), (3, 1000, 70), (4, 3, 139)])
def test_batched_param_draw_is_bit_identical_to_the_per_image_draw(seed, warm, n):
    """data.draw_augment_params_batch generates the MT19937 outputs itself (torch.get/set_rng_state): same parameter
    rows AND same generator state afterwards as n per-image draws, across state regenerations (624 outputs)."""
    from neuron_gan_b200 import data
    torch.manual_seed(seed)
    torch.rand(warm)                               # start somewhere inside the 624-word state
    start = torch.get_rng_state()
    slow = np.stack([data.draw_augment_params(96, 0.05) for _ in range(n)])
    after_slow, next_slow = torch.get_rng_state(), torch.randn(4)
    torch.set_rng_state(start)
    fast = data.draw_augment_params_batch(n, 96, 0.05)
    assert data._fast_draws_ok is True             # the fast path is live on this torch build
    assert torch.equal(torch.get_rng_state(), after_slow)
    assert np.array_equal(fast, slow)
    assert torch.equal(torch.randn(4), next_slow)  # latent draws that follow see the same stream


@pytest.mark.parametrize('canvas,crop,size', [(96, 64, 64), (96, 64, 16), (97, 64, 32), (75, 50, 50), (75, 50, 20)])
def test_oracle_matches_torchvision_live(canvas, crop, size):
    """The restatement against torchvision itself (when importable) on geometries the golden fixture does not hold:
    odd canvas - crop (the crop window and its vertical mirror differ by a row) and non power-of-two sizes."""
    tv = pytest.importorskip('torchvision')
    T = tv.transforms
    tr = [T.RandomAffine(degrees=180, translate=(0.05, 0.05), fill=0), T.RandomVerticalFlip(),
          T.ColorJitter(brightness=0.25, contrast=0.25), T.CenterCrop(size=crop), lambda t: t.mul(2).add(-1)]
    if size < crop:
        tr.append(T.Resize(size, antialias=True))
    tr = T.Compose(tr)
    g = torch.Generator().manual_seed(canvas * 1000 + size)
    for k in range(6):
        img = torch.rand(1, 1, canvas, canvas, generator=g)
        torch.manual_seed(50 + k)
        want = tr(img.clone())[0, 0].numpy()
        torch.manual_seed(50 + k)
        p = do.draw_params(canvas, 0.05)
        got = do.transform_image(img[0, 0].numpy(), p, crop, size)
        # nearest resampling: a source coordinate that lands on a half-integer to within float rounding may pick the
        # other neighbour (seen here: ix = 41.5 exactly in one evaluation order); allow 1 source pixel in 4000 to do
        # so, each of which reaches at most 2x2 antialiased outputs
        d = np.abs(got - want)
        allowed = 4 * int(np.ceil(2.5e-4 * crop * crop))
        assert (d > 1e-5).sum() <= allowed and np.median(d) <= 2e-6, (k, d.max(), (d > 1e-5).sum(), allowed)
