"""Smallest run that touches every kernel of the step (128x128 fade-in first: the vectorised image kernels and the
row-blocked resampling kernels need rows >= 128; then 32x32 fade-in, 16x16 and 64x64 stable; batch 2), for
`compute-sanitizer --tool memcheck python scripts/sanitize_small.py` on the B200 box."""
import sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200.train_step import TrainStep, build_networks
from oracle import pggan_oracle as O     # synthetic images only

for res, alpha in ((128, 0.5), (32, 0.5), (16, 1.0), (64, 1.0)):
    G, D = build_networks(res, alpha, seed=1, device='cuda')
    step = TrainStep(G, D, use_graph=False)
    for i in range(2):
        stats = step(O.synthetic_images(2, res, seed=i).cuda()).cpu()
        assert torch.isfinite(stats).all(), stats
    print(res, alpha, [round(float(v), 4) for v in stats], flush=True)
print('ok')
