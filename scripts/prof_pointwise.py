"""Short driver for `ncu --set full`: the pointwise / resampling kernels that run furthest below HBM bandwidth, once
each at the shapes of the 512x512 iteration (FromImage backward at 256x256, ToImage backward at 512x512, the upsample
adjoint 512 -> 256, the upsample 256 -> 512), rotating inputs so nothing is L2-resident."""
import sys, torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o
B, C = 16, 16
hi = [o.nchw_to_c8(torch.randn(B, C, 512, 512, device='cuda')) for _ in range(3)]
lo = [o.nchw_to_c8(torch.randn(B, C, 256, 256, device='cuda')) for _ in range(3)]
r_hi = torch.rand(B, 512, 512, device='cuda') + 0.5
r_lo = torch.rand(B, 256, 256, device='cuda') + 0.5
img_hi = torch.rand(B, 512, 512, device='cuda')
img_lo = torch.rand(B, 256, 256, device='cuda')
gimg = torch.zeros(B, 256, 256, device='cuda')
w = torch.randn(C, device='cuda')
gw, gb = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
for i in range(3):
    o.fromim_bwd(lo[i % 3], img_lo, w, gw, gb, g_img=gimg, accumulate=False)
    o.fromim_bwd(lo[(i + 1) % 3], img_lo, w, gw, gb)
    o.toim_bwd(img_hi, img_hi, hi[i % 3], r_hi, w, gw)
    o.up2_bwd_pn_bwd(hi[(i + 1) % 3], lo[i % 3], r_lo)
    o.upsample2x(lo[(i + 2) % 3])
torch.cuda.synchronize()
print('ok')
