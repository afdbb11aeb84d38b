"""Small driver for ncu: a few launches of the dominant high-resolution kernels (16->16 channels, 512x512,
16 images): fused conv forward, fused dgrad+PixelNorm backward, weight gradient."""
import math
import sys

import torch

sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o  # noqa: E402

B, C, R = 16, 16, 512
s = math.sqrt(2 / 1.04) / math.sqrt(C * 9)
x = o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda'))
w = torch.randn(C, C, 3, 3, device='cuda') * s
w_fwd, w_dg = o.prep_conv_weight(w)
dw = torch.zeros_like(w)
for _ in range(3):
    y, r = o.conv3x3_fwd(x, w_fwd, None, s, 0.2, C)
    ga, _ = o.conv3x3_dgrad_pn(y, w_dg, s, 0.2, x, r)
    o.conv3x3_wgrad(x, ga, s, dw)
    up = o.upsample2x(o.avgpool2(x))
torch.cuda.synchronize()
print('ok', float(y.float().abs().mean()), float(dw.abs().mean()))
