"""Phase timeline of the wgrad kernel (globaltimer stamps per CTA): launch->prologue->first tile->loop->flush."""
import ctypes, sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o, _lib
lib = _lib.load()
lib.ngan_debug_wgrad_trace.argtypes = [ctypes.c_void_p]
shapes = [(16, 16, 16, 512, 512), (16, 16, 16, 256, 256), (16, 32, 32, 128, 128), (16, 64, 32, 64, 64), (16, 128, 64, 32, 32), (16, 128, 128, 16, 16)]
for B, cin, cout, H, W in shapes:
    x = torch.randn(B, cin // 8, H, W, 8, device='cuda').bfloat16()
    g = torch.randn(B, cout // 8, H, W, 8, device='cuda').bfloat16()
    dw = torch.zeros(cout, cin, 3, 3, device='cuda')
    buf = torch.zeros(4096 * 8, dtype=torch.int64, device='cuda')
    for _ in range(2):
        o.conv3x3_wgrad(x, g, 0.1, dw)
    torch.cuda.synchronize()
    lib.ngan_debug_wgrad_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o.conv3x3_wgrad(x, g, 0.1, dw)
    e1.record()
    torch.cuda.synchronize()
    lib.ngan_debug_wgrad_trace(None)
    t = buf.view(-1, 8).cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    rel = (t[:, :5] - t0) / 1e3
    print(f'B={B} {cin}->{cout}@{H}: {t.shape[0]} CTAs, event {e0.elapsed_time(e1)*1e3:.1f} us; per-phase END times (us after first CTA start), median / max over CTAs:')
    for k, name in enumerate(['start', 'prologue+pdl_wait', 'first tile landed', 'main loop done', 'flush done']):
        print(f'    {name:20s} {rel[:, k].median():7.2f} {rel[:, k].max():7.2f}')
