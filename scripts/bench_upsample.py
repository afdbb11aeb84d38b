"""Graph-timed A/B of the x2 bilinear upsample at the shapes of the merged generator pass (B = 48) and of inference
(B = 128).  usage (B200): python scripts/bench_upsample.py      (runs itself once per NGAN_UP2_ROWS setting)"""
import os
import subprocess
import sys

if 'NGAN_UP2_ROWS' not in os.environ:
    for rows in ('1', '2', '4'):
        subprocess.run([sys.executable, __file__], env=dict(os.environ, NGAN_UP2_ROWS=rows), check=True)
    sys.exit(0)

import torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o
from scripts.bench_elementwise_lib import timeit

tot = 0.0
for B, C, R in ((48, 16, 512), (48, 32, 256), (48, 32, 128), (48, 64, 64), (48, 128, 32), (128, 16, 512)):
    n = 6 if B * C * R * R * 2 < 3e8 else 3
    lo = [o.nchw_to_c8(torch.randn(B, C, R // 2, R // 2, device='cuda')) for _ in range(n)]
    ref = o.upsample2x(lo[0]).float()
    us = timeit(lambda i: o.upsample2x(lo[i % n]))
    nbytes = B * C * R * R * 2 * 1.25
    if B == 48:
        tot += us
    print(f'rows={os.environ["NGAN_UP2_ROWS"]} upsample2x {R // 2}->{R} B={B} C={C}: {us:7.1f} us {nbytes / us / 1e3:6.0f} GB/s '
          f'checksum {ref.double().sum().item():.6f} {ref.double().abs().sum().item():.3f}', flush=True)
print(f'rows={os.environ["NGAN_UP2_ROWS"]} sum over the B=48 shapes: {tot:.1f} us', flush=True)
