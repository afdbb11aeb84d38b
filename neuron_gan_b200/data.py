"""The reference's on-device image pipeline (`config.image_preprocessing = 'device'`, train.py:149-154) behind
the reference's own names: a dataset object holding the preloaded padded canvases with `image_size`,
`image_size_max`, `set_image_size()` (data/NeuronDataset.py:149-164), and `DatasetIterator(dataset, batch_size,
device)` (data/NeuronDataset.py:170-205) whose batches are produced by three CUDA launches (csrc/augment.cu) instead
of six torchvision transforms per image.

Same results for the same torch seed: the random parameters are drawn here, on the host, with the same CPU-generator
calls in the same order as torchvision's RandomAffine.get_params, RandomVerticalFlip.forward and
ColorJitter.get_params make when the reference runs `self.transforms(image)` image by image; the kernels then do the
pixel work.  Images are served in dataset order with a ragged last batch, like the reference (no shuffling there).

Loading image files, the Otsu noise statistics and the zero-pixel noise fill (data/NeuronDataset.py:60-110) are
load-time CPU work outside this path: build the canvases with the reference's loader (or any other) and hand them
over with `NeuronImages(canvases, ...)` / `NeuronImages.from_dataset(reference_dataset)`.
"""
import math

import numpy as np
import torch

from . import ops
from .utils import PinnedRing

PARAM_FLOATS = 16


def draw_augment_params(canvas: int, translate: float, degrees: float = 180.0, brightness: float = 0.25,
                        contrast: float = 0.25):
    """One image's parameter row (layout: include/ngan_b200.h, ngan_augment_batch), consuming torch's global CPU
    generator exactly like the reference's three random transforms (data/NeuronDataset.py:113-117)."""
    angle = float(torch.empty(1).uniform_(-float(degrees), float(degrees)).item())
    max_d = float(translate * canvas)
    tx = int(round(torch.empty(1).uniform_(-max_d, max_d).item()))
    ty = int(round(torch.empty(1).uniform_(-max_d, max_d).item()))
    flip = bool(torch.rand(1) < 0.5)
    fn_idx = torch.randperm(4).tolist()
    b = float(torch.empty(1).uniform_(1 - brightness, 1 + brightness))
    c = float(torch.empty(1).uniform_(1 - contrast, 1 + contrast))
    # inverse of "rotate by angle about the centre, then translate" (torchvision _get_inverse_affine_matrix with
    # centre (0, 0), scale 1, no shear), computed in Python doubles like torchvision does
    rot = math.radians(angle)
    cs, sn = math.cos(rot), math.sin(rot)
    m = [cs, sn, 0.0, -sn, cs, 0.0]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    row = np.zeros(PARAM_FLOATS, dtype=np.float32)
    row[0:6] = np.asarray(m, dtype=np.float32) / np.float32(0.5 * canvas)     # _gen_affine_grid's rescaled theta
    row[6] = 1.0 if flip else 0.0
    row[7], row[8], row[9] = b, c, np.float32(1.0 - c)
    row[10] = 0.0 if fn_idx.index(0) < fn_idx.index(1) else 1.0               # brightness before contrast?
    return row


def identity_params():
    row = np.zeros(PARAM_FLOATS, dtype=np.float32)
    row[11] = 1.0
    return row


def aa_taps(in_size: int, out_size: int):
    """Antialiased-bilinear filter taps of ATen's _compute_indices_weights_aa (what Resize(antialias=True) uses),
    with ATen's mix of float and double intermediates.  Returns (first [R] i32, count [R] i32, weight [R, T] f32)."""
    scale = np.float32(np.float32(in_size) / np.float32(out_size))
    support = np.float32(scale) if scale >= 1.0 else np.float32(1.0)
    inv = np.float32(1.0 / scale) if scale >= 1.0 else np.float32(1.0)
    first = np.zeros(out_size, dtype=np.int32)
    count = np.zeros(out_size, dtype=np.int32)
    rows = []
    for i in range(out_size):
        center = np.float32(float(scale) * (i + 0.5))
        lo = max(int(float(np.float32(center - support)) + 0.5), 0)
        n = min(int(float(np.float32(center + support)) + 0.5), in_size) - lo
        w = np.zeros(n, dtype=np.float32)
        for j in range(n):
            x = np.float32((float(np.float32(np.float32(j + lo) - center)) + 0.5) * float(inv))
            w[j] = max(np.float32(0.0), np.float32(1.0) - abs(x))
        total = np.float32(0.0)
        for j in range(n):
            total = np.float32(total + w[j])
        w = w / total
        while n > 1 and w[n - 1] == 0.0:          # taps of weight exactly 0 (e.g. in == out: [1, 0]) cost work only
            n -= 1
        while n > 1 and w[0] == 0.0:
            w, lo, n = w[1:], lo + 1, n - 1
        rows.append(w[:n])
        first[i], count[i] = lo, n
    weight = np.zeros((out_size, int(count.max())), dtype=np.float32)
    for i, w in enumerate(rows):
        weight[i, :len(w)] = w
    return first, count, weight


class NeuronImages:
    """What DatasetIterator needs of the reference's NeuronDataset: the preloaded canvases (each image padded by
    image_size // 4 on every side and noise-filled, data/NeuronDataset.py:73-75, 100-107) and the output size."""

    load_all = True

    def __init__(self, canvases, image_size: int, augmentations: bool = True, im_translation: float = 0.0):
        if isinstance(canvases, (list, tuple)):
            canvases = torch.stack([torch.as_tensor(c).reshape(c.shape[-2], c.shape[-1]) for c in canvases])
        canvases = torch.as_tensor(canvases, dtype=torch.float32)
        if canvases.dim() == 4:
            if canvases.shape[1] != 1:
                raise ValueError('single-channel images expected (N_colors = 1, configs/config.py:63)')
            canvases = canvases[:, 0]
        if canvases.dim() != 3 or canvases.shape[-1] != canvases.shape[-2]:
            raise ValueError('canvases must be [N, P, P] (or [N, 1, P, P]) square images')
        if canvases.shape[-1] < image_size:
            raise ValueError('canvases are smaller than image_size')
        self.canvases = canvases.contiguous()
        self.image_size = image_size
        self.image_size_max = image_size
        self.augmentations = bool(augmentations)
        self.im_translation = float(im_translation)

    @classmethod
    def from_dataset(cls, dataset, im_translation: float = 0.0, augmentations: bool = True):
        """From a loaded reference NeuronDataset (its `.images` list, data/NeuronDataset.py:107)."""
        if not getattr(dataset, 'load_all', True):
            raise Exception('On-device iteration is only possible when all images are loaded.')
        return cls(list(dataset.images), dataset.image_size_max, augmentations, im_translation)

    def set_image_size(self, size: int):
        assert size <= self.image_size_max, 'The image size ({}) must be < {}.'.format(size, self.image_size_max)
        self.image_size = size

    def __len__(self):
        return self.canvases.shape[0]


class DatasetIterator:
    """Mirror of the reference's DatasetIterator (data/NeuronDataset.py:170-205): iterating yields
    [b, 1, image_size, image_size] float32 batches in [-1, 1] on `device`, in dataset order, the last one ragged;
    the returned tensor is a view of one reused buffer.  `rank`/`world` (not in the reference, which is single-GPU):
    every rank draws the parameters of the whole global batch so the generator streams stay identical, and
    produces only its own rows [rank*b/world, (rank+1)*b/world)."""

    def __init__(self, dataset: NeuronImages, batch_size: int, device, rank: int = 0, world: int = 1):
        if not dataset.load_all:
            raise Exception('On-device iteration is only possible when all images are loaded.')
        self.dataset = dataset
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise ops._lib.NganError('neuron_gan_b200.data.DatasetIterator needs a CUDA device (no CPU fallback)')
        self.N_images = len(dataset)
        self.batch_size = batch_size
        self.rank, self.world = rank, world
        self.images = dataset.canvases.to(self.device)
        self.img_dtype = torch.float32
        self.output_size = None
        self.images_tensor = None
        self._taps = None
        self._ring_f = PinnedRing([(batch_size, PARAM_FLOATS)], torch.float32)
        self._ring_i = PinnedRing([(batch_size,)], torch.int32)
        self._params = torch.empty((batch_size, PARAM_FLOATS), dtype=torch.float32, device=self.device)
        self._index = torch.empty((batch_size,), dtype=torch.int32, device=self.device)
        self._workspace = None
        self.image_ind = 0

    def _prepare(self):
        size = self.dataset.image_size
        if self.output_size is None or self.output_size[-1] != size:
            self.output_size = (self.batch_size, 1, size, size)
            self.images_tensor = torch.zeros(self.output_size, device=self.device, dtype=self.img_dtype)
            first, count, weight = aa_taps(self.dataset.image_size_max, size)
            self._taps = tuple(torch.from_numpy(a).to(self.device) for a in (first, count, weight))

    def __iter__(self):
        self._prepare()
        self.image_ind = 0
        return self

    def __next__(self):
        if self.image_ind >= self.N_images:
            raise StopIteration
        ds = self.dataset
        P = self.images.shape[-1]
        n = min(self.batch_size, self.N_images - self.image_ind)
        kf, (host_p,) = self._ring_f.acquire()
        ki, (host_i,) = self._ring_i.acquire()
        rows = host_p.numpy()
        for i in range(n):          # the whole (global) batch: keeps the RNG stream that of the reference
            rows[i] = draw_augment_params(P, ds.im_translation) if ds.augmentations else identity_params()
            host_i[i] = self.image_ind + i
        self.image_ind += n
        lo, hi = self.rank * n // self.world, (self.rank + 1) * n // self.world
        if hi == lo:
            return self.images_tensor[:0]
        m = hi - lo
        self._params[:m].copy_(host_p[lo:hi], non_blocking=True)
        self._index[:m].copy_(host_i[lo:hi], non_blocking=True)
        self._ring_f.release(kf)
        self._ring_i.release(ki)
        need = ops._lib.call('ngan_augment_workspace_bytes', m, P, ds.image_size_max)
        if self._workspace is None or self._workspace.numel() * 4 < need:
            self._workspace = torch.empty((need + 3) // 4, dtype=torch.float32, device=self.device)
        out = self.images_tensor[:m]
        ops.augment_batch(self.images, self._index[:m], self._params[:m], *self._taps, out, ds.image_size_max,
                          self._workspace)
        return out
