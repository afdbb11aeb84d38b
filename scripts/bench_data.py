"""the image pipeline.  For the BASELINE geometry (512 px images on 768 px canvases, 16 per
batch) at every phase's output size, prints one JSON line with
  * the augment kernels' time per batch (CUDA events around a captured graph of 20 batches, fresh parameters
    per batch) and the host time to draw a batch's parameters,
  * the same batches through the transform list of the reference's dataset (torchvision on CPU tensors: the
    reference's `image_preprocessing = 'cpu'` path without DataLoader workers), timed on the host,
  * with `--train`: images/s of the training step fed by the device iterator, end to end.
usage (B200): python scripts/bench_data.py [--train]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from neuron_gan_b200 import data, ops

B, P, CROP = 16, 768, 512
cv = torch.rand(64, P, P, generator=torch.Generator().manual_seed(0))
ds = data.NeuronImages(cv, CROP, True, 0.05)


def reference_cpu_transforms(size):
    import torchvision
    T = torchvision.transforms
    tr = [T.RandomAffine(degrees=180, translate=(0.05, 0.05), fill=0), T.RandomVerticalFlip(),
          T.ColorJitter(brightness=0.25, contrast=0.25), T.CenterCrop(size=CROP),
          lambda t: t.mul_(2).add_(-1)]
    if size < CROP:
        tr.append(T.Resize(size, antialias=True))
    return T.Compose(tr)


if '--once' in sys.argv:            # one eager batch per size, for ncu
    for size in (512, 256, 128, 32):
        ds.set_image_size(size)
        for x in data.DatasetIterator(ds, B, 'cuda'):
            break
        torch.cuda.synchronize()
    sys.exit(0)

for size in (512, 256, 128, 64, 32, 16):
    ds.set_image_size(size)
    it = data.DatasetIterator(ds, B, 'cuda')
    it._prepare()
    n_rep = 20
    t0 = time.perf_counter()
    tables = torch.from_numpy(np.stack([data.draw_augment_params_batch(B, P, 0.05) for _ in range(n_rep)])).cuda()
    host_ms = (time.perf_counter() - t0) / n_rep * 1e3
    index = torch.arange(B, dtype=torch.int32, device='cuda')
    out = torch.empty(B, 1, size, size, device='cuda')
    ws = torch.empty(ops._lib.call('ngan_augment_workspace_bytes', B, P, CROP) // 4 + 4, device='cuda')
    idx = [index + k * B for k in range(4)]
    run = lambda: [ops.augment_batch(it.images, idx[k % 4], tables[k], *it._taps, out, CROP, ws) for k in range(n_rep)]
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / n_rep * 1e3
    tr = reference_cpu_transforms(size)
    t0 = time.perf_counter()
    for k in range(B):
        tr(cv[k:k + 1].clone())
    cpu_ms = (time.perf_counter() - t0) * 1e3
    line = {'workload': f'augment_{CROP}_to_{size}_b{B}', 'kernels_us_per_batch': round(dev_us, 1),
            'algorithmic_bytes': B * (P * P + size * size) * 4,
            'GBps': round(B * (P * P + size * size) * 4 / dev_us / 1e3, 1),
            'host_param_draw_ms_per_batch': round(host_ms, 3),
            'device_images_per_s': round(B / max(dev_us * 1e-6, host_ms * 1e-3), 0),
            'torchvision_cpu_ms_per_batch': round(cpu_ms, 1), 'torchvision_cpu_images_per_s': round(B / cpu_ms * 1e3, 1),
            'cpu_threads': torch.get_num_threads()}
    print(json.dumps(line), flush=True)

if '--train' in sys.argv:
    from neuron_gan_b200.train_step import TrainStep, build_networks
    G, D = build_networks(512, 1.0, seed=1, device='cuda')
    step = TrainStep(G, D)
    ds.set_image_size(512)
    it = data.DatasetIterator(ds, B, 'cuda')
    for epoch in range(3):                      # 4 batches per epoch; first epochs warm up / capture the graph
        for x in it:
            step(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for epoch in range(10):
        for x in it:
            stats = step(x)
            n += x.shape[0]
    stats.cpu()
    dt = time.perf_counter() - t0
    print(json.dumps({'workload': 'train_512_b16_fed_by_device_iterator', 'images_per_s': round(n / dt, 1),
                      'ms_per_iteration': round(dt / (n / B) * 1e3, 3)}), flush=True)
