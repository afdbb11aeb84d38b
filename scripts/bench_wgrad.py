"""Graph-timed micro-benchmark of ngan_conv3x3_wgrad over the shapes of the 512x512 step (16 images).
    python scripts/bench_wgrad.py [lib.so] [old|new]
`old` drives the round-1 ABI (atomics, no workspace) so that a library built from an earlier commit can be timed
beside the current one on the same box."""
import ctypes
import sys

import torch

lib_path = sys.argv[1] if len(sys.argv) > 1 else 'neuron_gan_b200/libngan_b200.so'
abi = sys.argv[2] if len(sys.argv) > 2 else 'new'
lib = ctypes.CDLL(lib_path)
vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
if abi == 'new':
    lib.ngan_conv3x3_wgrad.argtypes = [vp, vp, f32, vp, i32, vp, i32, i32, i32, i32, i32, vp]
    lib.ngan_conv3x3_wgrad_workspace_bytes.restype = ctypes.c_longlong
    lib.ngan_conv3x3_wgrad_workspace_bytes.argtypes = [i32] * 5
else:
    lib.ngan_conv3x3_wgrad.argtypes = [vp, vp, f32, vp, i32, i32, i32, i32, i32, vp]
lib.ngan_last_error.restype = ctypes.c_char_p

SHAPES = [(16, 16, 16, 512, 512), (16, 16, 16, 256, 256), (32, 16, 16, 256, 256), (16, 32, 16, 256, 256),
          (16, 32, 32, 128, 128), (16, 16, 32, 128, 128), (16, 32, 32, 64, 64), (16, 64, 32, 64, 64),
          (16, 64, 64, 32, 32), (16, 128, 64, 32, 32), (16, 128, 128, 16, 16), (32, 128, 128, 16, 16)]
n_buf, iters = 6, 20
for B, cin, cout, H, W in SHAPES:
    xs = [torch.randn(B, cin // 8, H, W, 8, device='cuda').bfloat16() for _ in range(n_buf)]
    gs = [torch.randn(B, cout // 8, H, W, 8, device='cuda').bfloat16() for _ in range(n_buf)]
    dw = torch.zeros(cout, cin, 3, 3, device='cuda')
    ws = None
    if abi == 'new':
        ws = torch.empty(max(4, lib.ngan_conv3x3_wgrad_workspace_bytes(B, cin, cout, H, W)) // 4, device='cuda')

    def call(i):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        x, g = xs[i % n_buf], gs[i % n_buf]
        if abi == 'new':
            rc = lib.ngan_conv3x3_wgrad(x.data_ptr(), g.data_ptr(), 0.1, dw.data_ptr(), 1, ws.data_ptr(), B, cin, cout,
                                        H, W, st)
        else:
            rc = lib.ngan_conv3x3_wgrad(x.data_ptr(), g.data_ptr(), 0.1, dw.data_ptr(), B, cin, cout, H, W, st)
        assert rc == 0, lib.ngan_last_error()

    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            call(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    nbytes = B * H * W * (cin + cout) * 2
    print(f'{abi} wgrad B={B} {cin}->{cout} @{H}: {best:6.1f} us  {nbytes / best / 1e3:6.0f} GB/s  '
          f'{2 * 9 * cin * cout * B * H * W / best / 1e6:6.1f} TFLOP/s', flush=True)
