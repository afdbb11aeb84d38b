import ctypes, math, os, sys, torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o, _lib
B, C, R = 16, 16, 512
s = math.sqrt(2 / 1.04) / math.sqrt(C * 9)
x = o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda'))
w = torch.randn(C, C, 3, 3, device='cuda') * s
w_fwd, w_dg = o.prep_conv_weight(w)
for _ in range(3):
    y, r = o.conv3x3_fwd(x, w_fwd, None, s, 0.2, C)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 256)()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.ngan_debug_conv_trace(buf) == 0
t = [[buf[i * 8 + k] for k in range(8)] for i in range(32)]
t0 = t[0][0]
print('tile: mma[start waitEmpty waitFull issued lastEpiWarpDone]  epi(warp 3)[start gotFull done]   (cycles since first)')
for i in range(24):
    print(i, [v - t0 for v in t[i][:5]], [v - t0 for v in t[i][5:]])
