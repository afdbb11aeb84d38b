"""Short driver for `ncu --set full`: the step's heaviest conv launches in isolation (512x512, 16 -> 16 channels,
batch 32 = the critic batch [real; fake] / the two detached generator passes), forward, data gradient and weight
gradient, three rotating inputs so nothing is L2-resident."""
import math, sys, torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o
B, C, R = 32, 16, 512
s = math.sqrt(2 / 1.04) / math.sqrt(C * 9)
xs = [o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda')) for _ in range(3)]
w = torch.randn(C, C, 3, 3, device='cuda') * s
w_fwd, w_dg = o.prep_conv_weight(w)
dw = torch.zeros_like(w)
for i in range(4):
    y, r = o.conv3x3_fwd(xs[i % 3], w_fwd, None, s, 0.2, C)
    g = o.conv3x3_dgrad(xs[(i + 1) % 3], w_dg, s, C)
    o.conv3x3_wgrad(xs[i % 3], xs[(i + 2) % 3], s, dw)
torch.cuda.synchronize()
print('ok')
