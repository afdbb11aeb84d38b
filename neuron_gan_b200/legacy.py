"""Importable placeholders for the reference's four legacy network classes (reference models.py:622-790: BatchNorm
DCGAN and weight-clipping WGAN nets).  `train.py:108` does `from models import *` and the reference's `__all__`
(models.py:10-12) lists them, so the names must exist for the unmodified script to import; they are a different
algorithm (OUT OF SCOPE) and are not rebuilt: constructing one raises."""
import torch.nn as nn

__all__ = ['Generator_dcgan', 'Discriminator_dcgan', 'Generator_wgan', 'Discriminator_wgan']


class _NotBuilt(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(
            f'{type(self).__name__} (reference models.py:622-790) is outside the B200 build: only the progressive-'
            f'growing WGAN-GP path (Generator_PG / Discriminator_PG, config.pggan = True) runs on neuron_gan_b200')


class Generator_dcgan(_NotBuilt):
    pass


class Discriminator_dcgan(_NotBuilt):
    pass


class Generator_wgan(_NotBuilt):
    pass


class Discriminator_wgan(_NotBuilt):
    pass
