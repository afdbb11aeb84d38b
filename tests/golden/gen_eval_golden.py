"""Generate tests/golden/eval_golden.pt: the UNMODIFIED reference's gen_samples (utils.py:346-355) on CPU for the eval
configuration (BASELINE config 5): seeded latents (seed = 0) through the 512x512 generator built from seed 1, and the
same at 64x64 during a fade-in.  Records z, image patches and per-image checksums.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_eval_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402


def main():
    _, _, ref_utils = ref_loader.load()
    out = {}
    for res, alpha, n in ((512, 1.0, 4), (64, 0.5, 6)):
        G, _ = ref_loader.build_nets(res, alpha)
        ref_utils.Latent_vecs_memo.clear()
        images, z = ref_utils.gen_samples(G, N_images=n, seed=0)
        out[f'r{res}_a{alpha}'] = {'n': n, 'z': z.clone(), 'patch': images[:, 0, :16, :16].clone(),
                                   'sum': images.double().sum(dim=(1, 2, 3)), 'abssum': images.double().abs().sum(dim=(1, 2, 3)),
                                   'shape': tuple(images.shape)}
        print(res, alpha, out[f'r{res}_a{alpha}']['sum'])
    path = os.path.join(HERE, 'eval_golden.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
