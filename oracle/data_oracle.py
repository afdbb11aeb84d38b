"""CPU oracle for neuron-gan's on-device image pipeline (`image_preprocessing = 'device'`).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``neuron_gan_b200``) may import, call or execute
this module; only ``tests/`` does, as the checker.

What it restates: what ``DatasetIterator.__next__`` (reference
data/NeuronDataset.py:170-205) applies to every preloaded, padded image -- the transform list built in
``NeuronDataset.__init__`` (data/NeuronDataset.py:112-126) and extended by ``set_image_size``
(data/NeuronDataset.py:149-164):

    RandomAffine(degrees=180, translate=(t, t), fill=0)   nearest-neighbour resampling
    RandomVerticalFlip()                                   p = 0.5
    ColorJitter(brightness=0.25, contrast=0.25)            random order of the two
    CenterCrop(image_size_max)
    Renormalize((-1, 1), (0, 1))                           x*2 - 1, data/NeuronDataset.py:24-43
    Resize(image_size, antialias=True)                     only while image_size < image_size_max

The arithmetic of these transforms lives in third-party dependencies of the reference, torchvision
(``requirements.txt``: ``torchvision==0.14.1``; this image has 0.26) and ATen's ``grid_sampler_2d`` /
``_upsample_bilinear2d_aa``.  The functions below restate those published algorithms in plain numpy loops /
vector expressions -- they do not call torchvision:

  * parameter draws: ``RandomAffine.get_params``, ``RandomVerticalFlip.forward``, ``ColorJitter.get_params``
    (torchvision/transforms/transforms.py) -- same torch CPU-RNG calls in the same order;
  * ``_get_inverse_affine_matrix`` (transforms/functional.py), ``_gen_affine_grid`` + ``grid_sample(mode='nearest',
    padding_mode='zeros', align_corners=False)`` (transforms/_functional_tensor.py);
  * ``_blend`` / ``adjust_brightness`` / ``adjust_contrast`` (transforms/_functional_tensor.py);
  * ``center_crop`` (transforms/functional.py);
  * ATen ``_compute_indices_weights_aa`` with the bilinear (triangle) filter
    (aten/src/ATen/native/cpu/UpSampleKernel.cpp).

Pinning: the reference has no tests for this path.  ``tests/golden/gen_data_golden.py`` runs the UNMODIFIED
reference ``NeuronDataset`` + ``DatasetIterator`` in the build container (skimage, absent from the image, stubbed
for the load-time Otsu threshold only) and commits its source canvases and output batches;
``tests/test_data_oracle.py`` checks this file against them.

Nearest-neighbour resampling is discontinuous: a source coordinate within float rounding of a half-integer can
select the neighbouring pixel under a different (equally valid) fp32 evaluation order of the same affine
expression.  Comparisons therefore allow a small FRACTION of differing pixels rather than a uniform bound; the
tests state it.
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ------------------------------------------------------------------------------------------- parameter draws
def draw_params(canvas: int, translate: float, degrees: float = 180.0, brightness: float = 0.25,
                contrast: float = 0.25, augmentations: bool = True) -> dict:
    """One image's random parameters, drawn from torch's global CPU generator exactly as the three random
    transforms do when the reference calls ``self.transforms(image)`` (data/NeuronDataset.py:199)."""
    if not augmentations:
        return dict(angle=0.0, tx=0, ty=0, flip=False, order=0, b=1.0, c=1.0, identity=True)
    angle = float(torch.empty(1).uniform_(-degrees, degrees).item())
    max_d = float(translate * canvas)
    tx = int(round(torch.empty(1).uniform_(-max_d, max_d).item()))
    ty = int(round(torch.empty(1).uniform_(-max_d, max_d).item()))
    flip = bool(torch.rand(1) < 0.5)
    fn_idx = torch.randperm(4).tolist()
    b = float(torch.empty(1).uniform_(1 - brightness, 1 + brightness))
    c = float(torch.empty(1).uniform_(1 - contrast, 1 + contrast))
    order = 0 if fn_idx.index(0) < fn_idx.index(1) else 1          # 0: brightness first, 1: contrast first
    return dict(angle=angle, tx=tx, ty=ty, flip=flip, order=order, b=b, c=c, identity=False)


def inverse_affine_matrix(angle: float, tx: float, ty: float) -> list:
    """``_get_inverse_affine_matrix(center=[0,0], angle, translate, scale=1, shear=[0,0])`` in Python floats:
    output pixel (x, y), centred, maps to source (m0 x + m1 y + m2, m3 x + m4 y + m5)."""
    rot = math.radians(angle)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    return m


# ------------------------------------------------------------------------------------------- pixel transforms
def affine_nearest(img: np.ndarray, matrix: list) -> np.ndarray:
    """Tensor ``affine`` with nearest resampling and zero fill on a square [P, P] fp32 canvas."""
    P = img.shape[0]
    f = np.float32
    theta = np.asarray(matrix, dtype=f).reshape(2, 3) / f(0.5 * P)             # rescaled_theta (transposed)
    xs = (np.arange(P, dtype=f) - f(P * 0.5) + f(0.5))                         # linspace(-P/2+.5, P/2-.5, P)
    gx = (xs[None, :] * theta[0, 0] + xs[:, None] * theta[0, 1]) + theta[0, 2]
    gy = (xs[None, :] * theta[1, 0] + xs[:, None] * theta[1, 1]) + theta[1, 2]
    ix = ((gx + f(1)) * f(P) - f(1)) / f(2)                                    # unnormalise, align_corners=False
    iy = ((gy + f(1)) * f(P) - f(1)) / f(2)
    jx = np.rint(ix).astype(np.int64)                                          # nearbyint: half to even
    jy = np.rint(iy).astype(np.int64)
    inside = (jx >= 0) & (jx < P) & (jy >= 0) & (jy < P)
    out = np.zeros_like(img)
    out[inside] = img[jy[inside], jx[inside]]
    return out


def blend(img: np.ndarray, other, ratio: float) -> np.ndarray:
    f = np.float32
    return np.clip(f(ratio) * img + f(1.0 - ratio) * other, f(0), f(1)).astype(f)


def color_jitter(img: np.ndarray, b: float, c: float, order: int) -> np.ndarray:
    for which in ((0, 1) if order == 0 else (1, 0)):
        if which == 0:
            img = blend(img, np.float32(0), b)
        else:
            img = blend(img, img.mean(dtype=np.float32), c)
    return img


def aa_weights(in_size: int, out_size: int):
    """Per output index: (first source index, normalised triangle-filter weights), ``_compute_indices_weights_aa``."""
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = scale if scale >= 1 else f(1)
    inv = f(1) / scale if scale >= 1 else f(1)
    table = []
    for i in range(out_size):
        center = scale * f(i + 0.5)
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        w = np.array([max(f(0), f(1) - abs((f(j + lo) - center + f(0.5)) * inv)) for j in range(n)], dtype=f)
        table.append((lo, w / w.sum(dtype=f)))
    return table


def resize_aa(img: np.ndarray, out_size: int) -> np.ndarray:
    """``interpolate(mode='bilinear', antialias=True, align_corners=False)`` of a square image, separable:
    horizontal pass, then vertical, as ATen's CPU kernel orders them."""
    n = img.shape[0]
    table = aa_weights(n, out_size)
    tmp = np.zeros((n, out_size), dtype=np.float32)
    for i, (lo, w) in enumerate(table):
        tmp[:, i] = (img[:, lo:lo + len(w)] * w[None, :]).sum(axis=1, dtype=np.float32)
    out = np.zeros((out_size, out_size), dtype=np.float32)
    for i, (lo, w) in enumerate(table):
        out[i, :] = (tmp[lo:lo + len(w), :] * w[:, None]).sum(axis=0, dtype=np.float32)
    return out


def transform_image(canvas_img: np.ndarray, p: dict, crop: int, out_size: int) -> np.ndarray:
    """The whole transform list on one [P, P] canvas -> [out_size, out_size] in [-1, 1]."""
    img = np.asarray(canvas_img, dtype=np.float32)
    if not p.get('identity', False):
        img = affine_nearest(img, inverse_affine_matrix(p['angle'], p['tx'], p['ty']))
        if p['flip']:
            img = img[::-1, :]
        img = color_jitter(img, p['b'], p['c'], p['order'])
    P = img.shape[0]
    top = int(round((P - crop) / 2.0))
    img = img[top:top + crop, top:top + crop]
    img = (img - np.float32(0)) / np.float32(1) * np.float32(2) + np.float32(-1)
    if out_size < crop:
        img = resize_aa(np.ascontiguousarray(img), out_size)
    return np.ascontiguousarray(img, dtype=np.float32)


def epoch_batches(canvases, batch_size: int, crop: int, out_size: int, translate: float,
                  augmentations: bool = True):
    """``DatasetIterator`` (data/NeuronDataset.py:170-205): dataset order, no shuffling, ragged last batch.
    Yields (batch [b, 1, R, R] float32, list of parameter dicts)."""
    P = canvases[0].shape[-1]
    for first in range(0, len(canvases), batch_size):
        rows, params = [], []
        for k in range(first, min(first + batch_size, len(canvases))):
            p = draw_params(P, translate, augmentations=augmentations)
            params.append(p)
            rows.append(transform_image(np.asarray(canvases[k]).reshape(P, P), p, crop, out_size))
        yield np.stack(rows)[:, None], params
