"""BASELINE config 5: eval.py's path -- generator-only inference through utils.gen_samples (reference utils.py:346-355,
eval.py:23-26) at 512x512, n = 20 ... 4096 samples, chunked.  Prints one JSON line per n.
usage (B200): python scripts/bench_eval.py [chunk]"""
import json
import sys
import time
import torch
sys.path.insert(0, '.')
from neuron_gan_b200.train_step import build_networks
from neuron_gan_b200.utils import gen_samples

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G, _ = build_networks(512, 1.0, seed=1, device='cuda')
G.train(False)
for n in (20, 64, 256, 1024, 4096):
    gen_samples(G, min(n, chunk), seed=0, chunk=chunk)            # warm-up (weight images, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    images, z = gen_samples(G, n, seed=0, chunk=chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert images.shape == (n, 1, 512, 512) and torch.isfinite(images).all()
    print(json.dumps({'workload': 'eval_generator_512', 'n': n, 'chunk': chunk, 'seconds': round(dt, 4),
                      'images_per_s': round(n / dt, 1), 'output': 'fp32 [n,1,512,512] on the device',
                      'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}), flush=True)
    del images
