"""Generate tests/golden/data_pipeline_golden.pt by running the UNMODIFIED reference image pipeline
(/root/reference/data/NeuronDataset.py: NeuronDataset + DatasetIterator) in the build container.

`skimage` is absent from the image; it is needed only for the load-time Otsu noise threshold
(data/NeuronDataset.py:90-93) and is stubbed with a fixed threshold.  The fixture keeps what the per-batch path
consumes and produces: the preloaded padded canvases (`dataset.images`), and for several output sizes the batches
`DatasetIterator` yields after `torch.manual_seed(seed)`.

    python tests/golden/gen_data_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get('NGAN_REFERENCE_DIR', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_data_module():
    sk, skf = types.ModuleType('skimage'), types.ModuleType('skimage.filters')
    skf.threshold_multiotsu = lambda a, classes=4: np.array([40.0, 90.0, 160.0])
    sk.filters = skf
    sys.modules.setdefault('skimage', sk)
    sys.modules.setdefault('skimage.filters', skf)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from data import NeuronDataset as ND
    return ND


def main():
    from PIL import Image
    ND = load_reference_data_module()
    S, N, B, T = 64, 5, 2, 0.05
    d = tempfile.mkdtemp()
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:S, 0:S]
    for i in range(N):
        # smooth structure + texture, with true zeros for the load-time noise fill
        a = 120 + 100 * np.sin(xx / (3.0 + i)) * np.cos(yy / (4.0 + i)) + 25 * rng.randn(S, S)
        a = np.clip(a, 0, 255).astype(np.uint8)
        a[rng.rand(S, S) < 0.2] = 0
        Image.fromarray(a).save(os.path.join(d, f'im{i}.png'))
    np.random.seed(5)
    ds = ND.NeuronDataset(d, image_size=S, augmentations=True, im_translation=T)
    canvases = torch.stack([im.clone() for im in ds.images])                 # [N, 1, P, P]
    it = ND.DatasetIterator(ds, batch_size=B, device=torch.device('cpu'))
    out = {'canvases': canvases, 'image_size_max': S, 'batch_size': B, 'translate': T, 'epochs': []}
    for size, seed in ((64, 11), (32, 12), (16, 13), (8, 14)):
        ds.set_image_size(size)
        torch.manual_seed(seed)
        batches = [b.clone() for b in it]
        out['epochs'].append({'size': size, 'seed': seed, 'batches': batches})
    # without augmentations: crop + renormalise + resize only
    ds2 = ND.NeuronDataset(d, image_size=S, augmentations=False)
    ds2.images = [im.clone() for im in canvases]
    it2 = ND.DatasetIterator(ds2, batch_size=B, device=torch.device('cpu'))
    ds2.set_image_size(16)
    out['plain16'] = [b.clone() for b in it2]
    torch.save(out, os.path.join(HERE, 'data_pipeline_golden.pt'))
    print('wrote', os.path.join(HERE, 'data_pipeline_golden.pt'),
          os.path.getsize(os.path.join(HERE, 'data_pipeline_golden.pt')), 'bytes')


if __name__ == '__main__':
    main()
