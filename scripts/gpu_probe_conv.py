"""Diagnostic probe for the tcgen05 conv kernel (run on the GPU box): prints error summaries for a few
structured inputs so that a wrong descriptor / layout assumption can be identified from one run."""
import math
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from neuron_gan_b200 import ops as o  # noqa: E402


def run(B, cin, cout, H, W, w, x, tag):
    w_fwd, w_dg = o.prep_conv_weight(w)
    ga = o.conv3x3_dgrad(o.nchw_to_c8(x), w_fwd, 1.0, cout)   # EPI_LINEAR on the fwd image = plain conv, scale 1
    torch.cuda.synchronize()
    got = o.c8_to_nchw(ga)
    ref = F.conv2d(x, w, padding=1)
    err = (got - ref).abs()
    print(f'[{tag}] B={B} {cin}->{cout} {H}x{W}: max|err|={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} '
          f'rel={((got - ref).norm() / ref.norm()).item():.4g} nan={torch.isnan(got).any().item()}')
    return got, ref


def main():
    torch.manual_seed(0)
    dev = 'cuda'
    # NOTE conv3x3_dgrad(ga, w_img, scale, cin): ga has `cout` channels -> output `cin`; called with the FWD image
    # it computes out[cin_arg] = conv(x[cout_arg]) so pass (x with C=cin) and ask for cout channels.
    for (B, cin, cout, H, W) in [(1, 16, 16, 16, 16), (1, 16, 16, 64, 64), (2, 32, 32, 32, 32), (1, 64, 64, 32, 32),
                                 (1, 128, 128, 16, 16), (1, 16, 32, 32, 32), (1, 32, 16, 32, 32)]:
        x = torch.randn(B, cin, H, W, device=dev).bfloat16().float()
        w = (torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)).bfloat16().float()
        try:
            got, ref = run(B, cin, cout, H, W, w, x, 'random')
        except Exception as e:  # noqa: BLE001
            print('FAILED', (B, cin, cout, H, W), e)
            return
        if (got - ref).abs().max() > 0.05 and cin == 16 and H == 16:
            # centre-tap identity: output should equal input
            w2 = torch.zeros_like(w)
            for c in range(min(cin, cout)):
                w2[c, c, 1, 1] = 1
            got2, ref2 = run(B, cin, cout, H, W, w2, x, 'identity')
            print(' identity: got[0,:4,0,:4]=', got2[0, :4, 0, :4].tolist())
            print(' identity: ref[0,:4,0,:4]=', ref2[0, :4, 0, :4].tolist())
            # one-hot input
            x3 = torch.zeros_like(x)
            x3[0, 3, 5, 7] = 1
            w3 = torch.zeros_like(w)
            for t in range(9):
                w3[t % cout, 3, t // 3, t % 3] = t + 1
            got3, ref3 = run(B, cin, cout, H, W, w3, x3, 'onehot')
            nz = got3.nonzero()
            print(' onehot got nonzeros (first 20):', [(tuple(i.tolist()), got3[tuple(i.tolist())].item()) for i in nz[:20]])
            nz = ref3.nonzero()
            print(' onehot ref nonzeros:', [(tuple(i.tolist()), ref3[tuple(i.tolist())].item()) for i in nz[:20]])


if __name__ == '__main__':
    main()
