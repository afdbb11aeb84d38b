"""Micro-benchmark of the conv kernels on one shape (CUDA events, L2 flushed by rotating buffers)."""
import math
import sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o

def run(B, C, R, n_buf=6, iters=30):
    s = math.sqrt(2 / 1.04) / math.sqrt(C * 9)
    xs = [o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda')) for _ in range(n_buf)]
    w = torch.randn(C, C, 3, 3, device='cuda') * s
    w_fwd, w_dg = o.prep_conv_weight(w)
    y, r = o.conv3x3_fwd(xs[0], w_fwd, None, s, 0.2, C)
    res = {}
    def timeit(name, fn, nbytes):
        # the calls are captured into one CUDA graph and replayed: small kernels would otherwise be timed at the
        # host's launch rate (~20 us per Python call), not the GPU's
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        res[name] = (us, nbytes / us / 1e3)
    px = B * R * R
    timeit('fwd', lambda i: o.conv3x3_fwd(xs[i % n_buf], w_fwd, None, s, 0.2, C), px * (4 * C + 4))
    timeit('dgrad', lambda i: o.conv3x3_dgrad(xs[i % n_buf], w_dg, s, C), px * 4 * C)
    timeit('dgrad_pn', lambda i: o.conv3x3_dgrad_pn(xs[i % n_buf], w_dg, s, 0.2, xs[(i + 1) % n_buf], r), px * (6 * C + 4))
    timeit('dbl', lambda i: o.conv3x3_dbl(xs[i % n_buf], w_fwd, s, 0.2, xs[(i + 1) % n_buf], r, xs[(i + 2) % n_buf]), px * (10 * C + 4))
    dw = torch.zeros_like(w)
    timeit('wgrad', lambda i: o.conv3x3_wgrad(xs[i % n_buf], xs[(i + 1) % n_buf], s, dw), px * 4 * C)
    timeit('upsample', lambda i: o.upsample2x(xs[i % n_buf][:, :, :R // 2, :R // 2].contiguous()) if False else o.avgpool2(xs[i % n_buf]), px * 2 * C * 1.25)
    print(f'B={B} C={C} R={R}: ' + '  '.join(f'{k} {v[0]:.1f}us {v[1]:.0f}GB/s' for k, v in res.items()), flush=True)

if __name__ == '__main__':
    cfgs = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(16, 16, 512), (32, 16, 256), (16, 32, 128), (16, 64, 32)]
    for c in cfgs:
        run(*c)
