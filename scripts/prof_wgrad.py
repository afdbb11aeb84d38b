"""Three eager launches of ngan_conv3x3_wgrad on each of a few shapes (for `ncu -k regex:wgrad`)."""
import sys
import torch
sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o

shapes = [(16, 16, 16, 512, 512), (16, 32, 32, 128, 128), (16, 64, 32, 64, 64), (16, 128, 64, 32, 32), (16, 128, 128, 16, 16)]
for B, cin, cout, H, W in shapes:
    x = torch.randn(B, cin // 8, H, W, 8, device='cuda').bfloat16()
    g = torch.randn(B, cout // 8, H, W, 8, device='cuda').bfloat16()
    dw = torch.zeros(cout, cin, 3, 3, device='cuda')
    for _ in range(3):
        o.conv3x3_wgrad(x, g, 0.1, dw)
    torch.cuda.synchronize()
print('ok')
