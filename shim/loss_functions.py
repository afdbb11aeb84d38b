"""Drop-in for the reference's `loss_functions` module (train.py:23)."""
from neuron_gan_b200.loss_functions import D_W_loss, D_grad_pen_loss, G_W_loss, similarity_loss  # noqa: F401
