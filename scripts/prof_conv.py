"""Short driver for `ncu --set full`: the step's heaviest launches in isolation at the bench shape (512x512, 16 -> 16
channels, batch 32 = the critic batch [real; fake] / the two detached generator passes): folded conv forward, data
gradient, data gradient + PixelNorm backward, double backward, weight gradient (+ its reduction), and the
tensor-core Linear-gradient kernel; three rotating inputs so nothing is L2-resident."""
import math, sys, torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o
B, C, R = 32, 16, 512
s = math.sqrt(2 / 1.04) / math.sqrt(C * 9)
xs = [o.nchw_to_c8(torch.randn(B, C, R, R, device='cuda')) for _ in range(3)]
w = torch.randn(C, C, 3, 3, device='cuda') * s
w_fwd, w_dg = o.prep_conv_weight(w)
dw = torch.zeros_like(w)
r = torch.rand(B, R, R, device='cuda') + 0.5
for i in range(3):
    y, _ = o.conv3x3_fwd(xs[i % 3], w_fwd, None, s, 0.2, C)
    g = o.conv3x3_dgrad(xs[(i + 1) % 3], w_dg, s, C)
    o.conv3x3_dgrad_pn(xs[(i + 1) % 3], w_dg, s, 0.2, xs[i % 3], r)
    o.conv3x3_dbl(xs[i % 3], w_fwd, s, 0.2, xs[(i + 1) % 3], r, xs[(i + 2) % 3])
    o.conv3x3_wgrad(xs[i % 3], xs[(i + 2) % 3], s, dw, accumulate=False)
# the generator's Linear gradient from its factors, global batch 128 (8 ranks x 16) as one all-gathered buffer
Cl, S, K, Bt = 128, 16, 512, 128
ga = torch.randn(Bt, Cl // 8, S, S, 8, device='cuda').bfloat16()
z = torch.randn(Bt, K, device='cuda')
g_out = torch.empty(Cl * S * S, K, device='cuda')
for i in range(3):
    o.linear_wgrad_factored(ga, z, K, Cl, S, 0.01, g_out, b_per_seg=16, n_seg=8, ga_seg_stride=16 * Cl * S * S * 2,
                            z_seg_stride=16 * K * 4)
torch.cuda.synchronize()
print('ok')
