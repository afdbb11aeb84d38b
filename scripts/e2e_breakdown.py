"""Which part of the end-to-end path costs what: device/host images x device/CPU draws, statistics read back every
iteration (pipelined one iteration behind), 512x512, 16 images.  usage (B200): python scripts/e2e_breakdown.py"""
import sys
import time
import torch
sys.path.insert(0, '.')
from neuron_gan_b200.train_step import TrainStep, build_networks
from neuron_gan_b200.utils import DevicePrefetcher
from oracle import pggan_oracle as O     # synthetic images only

res, B, n = 512, 16, 40
G, D = build_networks(res, 1.0, seed=1, device='cuda')
step = TrainStep(G, D)
host = [O.synthetic_images(B, res, seed=100 + i).pin_memory() for i in range(4)]
devx = [h.cuda() for h in host]
draws = [step.draw(B, 'cuda') for _ in range(4)]
for i in range(4):
    step(devx[i % 4], draws[i % 4])
stat_bufs = [torch.empty(5).pin_memory() for _ in range(2)]


def run(host_images, cpu_draws, readback, iters):
    src = DevicePrefetcher(lambda: (host[j % 4] for j in range(iters)), 'cuda', gate=lambda: step.inputs_loaded) if host_images else \
        (devx[j % 4] for j in range(iters))
    pending = None
    for i, x in enumerate(src):
        s = step(x, None if cpu_draws else draws[i % 4])
        if readback:
            hb = stat_bufs[i % 2]
            hb.copy_(s, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending.synchronize()
            pending = ev
    torch.cuda.synchronize()


for host_images, cpu_draws, readback in ((0, 0, 0), (0, 0, 1), (0, 1, 1), (1, 0, 1), (1, 1, 1)):
    run(host_images, cpu_draws, readback, 4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(host_images, cpu_draws, readback, n)
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f'host_images={host_images} cpu_draws={cpu_draws} readback={readback}: {dt:.3f} ms / iteration', flush=True)
# host-side cost of the pieces
t0 = time.perf_counter()
for _ in range(50):
    step.draw_host(B)
print(f'draw_host: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms')
t0 = time.perf_counter()
for _ in range(50):
    step.opt_d.advance(); step.opt_g.advance()
torch.cuda.synchronize()
print(f'adam advance x2: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms')
