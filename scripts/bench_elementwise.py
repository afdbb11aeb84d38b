"""Graph-timed micro-benchmark of the pointwise / resampling kernels at the shapes of the 512x512 iteration.
usage (B200): python scripts/bench_elementwise.py"""
import sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o


def timeit(fn, iters=20):
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
    return e0.elapsed_time(e1) / iters * 1e3


def c8(B, C, R, n=4):
    return [o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda')) for _ in range(n)]


def report(name, us, nbytes):
    print(f'{name:44s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s', flush=True)


for B, C, R in ((16, 16, 512), (32, 16, 512), (16, 16, 256), (16, 32, 128), (16, 128, 16)):
    px = B * R * R
    xs = c8(B, C, R)
    lo = c8(B, C, R // 2)
    r = torch.rand(B, R, R, device='cuda') + 0.5
    img = torch.rand(B, R, R, device='cuda')
    w = torch.randn(C, device='cuda')
    bias = torch.randn(C, device='cuda')
    gw, gb = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    tag = f'B={B} C={C} R={R}'
    report(f'upsample2x ({R // 2}->{R}) {tag}', timeit(lambda i: o.upsample2x(lo[i % 4])), px * 2 * C * 1.25)
    report(f'avgpool2 ({R}->{R // 2}) {tag}', timeit(lambda i: o.avgpool2(xs[i % 4])), px * 2 * C * 1.25)
    report(f'pn_bwd {tag}', timeit(lambda i: o.pn_bwd(xs[i % 4], xs[(i + 1) % 4], r)), px * (6 * C + 4))
    report(f'pn_bwd unpool+gy {tag}', timeit(lambda i: o.pn_bwd(lo[i % 4], xs[(i + 1) % 4], r, 0.25, True, None, True)),
           px * (6.5 * C + 4))
    if R >= 32:
        rl = torch.rand(B, R // 2, R // 2, device='cuda') + 0.5
        report(f'up2_bwd_pn_bwd ({R}->{R // 2}) {tag}', timeit(lambda i: o.up2_bwd_pn_bwd(xs[i % 4], lo[(i + 1) % 4], rl)),
               px * 2 * C * 1.5 + px)
    report(f'toim_fwd {tag}', timeit(lambda i: o.toim_fwd(xs[i % 4], w)), px * (2 * C + 4))
    report(f'toim_bwd {tag}', timeit(lambda i: o.toim_bwd(img, img, xs[i % 4], r, w, gw)), px * (4 * C + 12))
    report(f'fromim_fwd {tag}', timeit(lambda i: o.fromim_fwd(img, w, bias)), px * (2 * C + 4))
    report(f'fromim_bwd {tag}', timeit(lambda i: o.fromim_bwd(xs[i % 4], img, w, gw, gb, g_img=img, accumulate=True)),
           px * (2 * C + 12))
