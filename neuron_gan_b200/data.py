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


# ---- the same draws for a whole batch at once -------------------------------------------------------------------
# draw_augment_params costs ~23 us per image in Python/torch call overhead (seven generator calls, like torchvision):
# 1.5 ms for a 64-image batch, twice the 16x16 training iteration it feeds.  torch's CPU generator is a plain MT19937
# whose state torch.get_rng_state() exposes; every draw above consumes exactly one 32-bit output (uniform_ on float32
# keeps its low 24 bits, randperm(4) takes three outputs modulo 4, 3, 2).  The batch version reads the state, produces
# the 9 outputs per image itself (vectorised twist + tempering), writes the advanced state back and applies the same
# transformations -- bit-identical values and generator state, checked once per process against the per-image path
# (any mismatch, e.g. a different state layout in another torch build, switches the fast path off).
_MT_N = 624
_U, _L, _A = np.uint32(0x80000000), np.uint32(0x7fffffff), np.uint32(0x9908b0df)
_fast_draws_ok = None


def _mt_twist(key):
    def run(lo, hi, src):
        y = (key[lo:hi] & _U) | (key[lo + 1:hi + 1] & _L)
        key[lo:hi] = key[src:src + hi - lo] ^ (y >> np.uint32(1)) ^ ((y & np.uint32(1)) * _A)
    run(0, 227, 397)            # needs the old key[397:624]
    run(227, 454, 0)            # needs the new key[0:227]
    run(454, 623, 227)
    y = (key[623] & _U) | (key[0] & _L)
    key[623] = key[396] ^ (y >> np.uint32(1)) ^ ((y & np.uint32(1)) * _A)


def _mt_temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000))
    return y ^ (y >> np.uint32(18))


def _take_raw(n):
    """The next n 32-bit outputs of torch's global CPU generator (which is advanced accordingly).  State layout
    (ATen CPUGeneratorImplStateLegacy): u64 seed | i32 left | i32 seeded | u64 next | u64 state[624] | normal cache."""
    st = torch.get_rng_state()
    a = st.numpy()
    if a.size < 24 + 8 * _MT_N:
        raise RuntimeError('unexpected CPU generator state size')
    left = int(a[8:12].view(np.int32)[0])
    words = a[24:24 + 8 * _MT_N].view(np.uint64)
    key = words.astype(np.uint32)
    pos = _MT_N if left == 1 else int(a[16:24].view(np.uint64)[0])
    if not (0 <= pos <= _MT_N and (left == 1 or left == _MT_N + 1 - pos)):
        raise RuntimeError('unexpected CPU generator state')
    out = np.empty(n, dtype=np.uint32)
    done = 0
    while done < n:
        if pos == _MT_N:
            _mt_twist(key)
            pos = 0
        m = min(n - done, _MT_N - pos)
        out[done:done + m] = _mt_temper(key[pos:pos + m])
        done += m
        pos += m
    words[:] = key
    a[8:12].view(np.int32)[0] = _MT_N + 1 - pos
    a[16:24].view(np.uint64)[0] = pos
    torch.set_rng_state(st)
    return out


def _fast_batch(n, canvas, translate, degrees, brightness, contrast):
    raw = _take_raw(9 * n).reshape(n, 9)
    u = (raw & np.uint32(0xFFFFFF)).astype(np.float64) * 2.0 ** -24

    def uniform(col, lo, hi):         # ATen uniform_real_distribution<float>: double math on float32 bounds
        lo, hi = np.float32(lo), np.float32(hi)
        return (u[:, col] * np.float64(hi - lo) + np.float64(lo)).astype(np.float32)

    max_d = float(translate * canvas)
    angle = uniform(0, -float(degrees), float(degrees))
    txs, tys = uniform(1, -max_d, max_d), uniform(2, -max_d, max_d)
    flip = u[:, 3].astype(np.float32) < np.float32(0.5)
    bs = uniform(7, 1 - brightness, 1 + brightness)
    cs = uniform(8, 1 - contrast, 1 + contrast)
    # randperm_cpu(4): swap k with k + random() % (4 - k), k = 0, 1, 2; only "brightness before contrast?" is used
    order = _PERM_ORDER[raw[:, 4] % np.uint32(4), raw[:, 5] % np.uint32(3), raw[:, 6] % np.uint32(2)]
    tx = np.rint(txs.astype(np.float64))          # Python round(): half to even, like rint
    ty = np.rint(tys.astype(np.float64))
    rot = [math.radians(float(v)) for v in angle]  # libm cos/sin on Python floats, exactly as the per-image path
    cs_ = np.array([math.cos(r) for r in rot], dtype=np.float64)
    sn = np.array([math.sin(r) for r in rot], dtype=np.float64)
    m = np.zeros((n, 6), dtype=np.float64)
    m[:, 0], m[:, 1], m[:, 3], m[:, 4] = cs_, sn, -sn, cs_
    m[:, 2] = m[:, 0] * (-tx) + m[:, 1] * (-ty)
    m[:, 5] = m[:, 3] * (-tx) + m[:, 4] * (-ty)
    rows = np.zeros((n, PARAM_FLOATS), dtype=np.float32)
    rows[:, 0:6] = m.astype(np.float32) / np.float32(0.5 * canvas)
    rows[:, 6] = flip
    rows[:, 7], rows[:, 8] = bs, cs
    rows[:, 9] = (1.0 - cs.astype(np.float64)).astype(np.float32)
    rows[:, 10] = order
    return rows


def _perm_order_table():
    t = np.zeros((4, 3, 2), dtype=np.float32)
    for z0 in range(4):
        for z1 in range(3):
            for z2 in range(2):
                perm = [0, 1, 2, 3]
                for k, z in enumerate((z0, z1, z2)):
                    perm[k], perm[k + z] = perm[k + z], perm[k]
                t[z0, z1, z2] = 0.0 if perm.index(0) < perm.index(1) else 1.0
    return t


_PERM_ORDER = _perm_order_table()


def draw_augment_params_batch(n, canvas, translate, degrees=180.0, brightness=0.25, contrast=0.25):
    """n parameter rows, bit-identical (values and generator state) to n calls of draw_augment_params."""
    global _fast_draws_ok
    args = (canvas, translate, degrees, brightness, contrast)
    if _fast_draws_ok is None:                  # one-time self check on the live generator state
        state = torch.get_rng_state()
        try:
            fast = _fast_batch(3, *args)
            after_fast = torch.get_rng_state()
            torch.set_rng_state(state)
            slow = np.stack([draw_augment_params(*args) for _ in range(3)])
            _fast_draws_ok = bool(np.array_equal(fast, slow) and torch.equal(after_fast, torch.get_rng_state()))
        except Exception:
            _fast_draws_ok = False
        torch.set_rng_state(state)
    if _fast_draws_ok:
        return _fast_batch(n, *args)
    return np.stack([draw_augment_params(*args) for _ in range(n)]) if n else np.zeros((0, PARAM_FLOATS), np.float32)


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
        # the whole (global) batch: keeps the RNG stream that of the reference
        rows[:n] = draw_augment_params_batch(n, P, ds.im_translation) if ds.augmentations else identity_params()
        host_i[:n] = torch.arange(self.image_ind, self.image_ind + n, dtype=torch.int32)
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
