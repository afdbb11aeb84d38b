"""neuron_gan_b200: B200-native (sm_100a) implementation of neuron-gan's progressive-growing WGAN-GP
training step behind the reference's Python module API.

    from neuron_gan_b200.models import Generator_PG, Discriminator_PG
    from neuron_gan_b200.loss_functions import D_W_loss, G_W_loss, D_grad_pen_loss
    from neuron_gan_b200.utils import sample_latent_vec, Checkpointer

All arithmetic on the hot path runs in hand-written CUDA kernels (csrc/) reached through the C ABI of
libngan_b200.so (include/ngan_b200.h).  There is no CPU fallback: using a network on a non-CUDA tensor, or
without the built library, raises.
"""
__version__ = '0.1.0'
