"""WGAN-GP losses with the reference's signatures and return conventions (reference loss_functions.py:7-74,
148-205), computed by the sm_100a kernels.

Differences from the reference that do not change results:
  * D_W_loss runs ONE critic pass over the concatenated [real; fake] batch (the critic has no cross-sample
    coupling) instead of two passes;
  * the generator forwards whose outputs the reference immediately detaches run without saving activations;
  * G_W_loss does not compute the critic's weight gradients, which the reference computes and discards
    (train.py:357 zeroes them before they are ever used);
  * D_grad_pen_loss takes an optional `epsilon` so a parity test can inject the interpolation draw; by
    default it draws torch.rand((B,1,1,1), device=device) exactly like loss_functions.py:170.
The NaN guards raise the same ValueErrors; set `check_nan = False` on a loss object to skip the host sync.
"""
import torch
import torch.nn as nn

from . import autograd_fns, ops
from .utils import sample_latent_vec, save_vars


class D_W_loss(nn.Module):
    """Wasserstein critic loss (+ drift). Returns (D_loss, score_real, score_fake)."""

    def __init__(self, generator_net, discriminator_net, drift_epsilon=0.0):
        super().__init__()
        self.generator_net = generator_net
        self.discriminator_net = discriminator_net
        self.drift_epsilon = drift_epsilon
        self.check_nan = True

    def forward(self, real_images):
        batch_size = real_images.size(0)
        device = real_images.device
        z = sample_latent_vec((batch_size, self.generator_net.latent_dim), device=device)
        with torch.no_grad():
            fake_images = self.generator_net(z)
        scores = self.discriminator_net(torch.cat([real_images.to(torch.float32), fake_images]))
        D_loss, score_real, score_fake = autograd_fns._WLossFn.apply(scores, batch_size, self.drift_epsilon)
        if self.check_nan:
            if torch.isnan(score_real):
                save_vars(locals())
                raise ValueError('Real loss is nan.')
            if torch.isnan(score_fake):
                save_vars(locals())
                raise ValueError('Fake loss is nan.')
        return D_loss, score_real, score_fake


class G_W_loss(nn.Module):
    """Wasserstein generator loss. Returns (G_loss, z_latent)."""

    def __init__(self, generator_net, discriminator_net):
        super().__init__()
        self.generator_net = generator_net
        self.discriminator_net = discriminator_net
        self.check_nan = True

    def forward(self, real_images_batch):
        batch_size = real_images_batch.size(0)
        device = real_images_batch.device
        z_latent = sample_latent_vec((batch_size, self.generator_net.latent_dim), device=device)
        fake_images = self.generator_net(z_latent)
        scores = autograd_fns.discriminator_forward(self.discriminator_net, fake_images, params_grad=False)
        G_loss = autograd_fns._GLossFn.apply(scores)
        if self.check_nan and torch.isnan(G_loss):
            save_vars(locals())
            raise ValueError('Generator loss is nan.')
        return G_loss, z_latent


class D_grad_pen_loss(nn.Module):
    """Gradient penalty Lambda * mean_b((||d D(x_hat)/d x_hat||_2 - 1)^2) (Gulrajani et al. 2017, Algorithm 1)."""

    def __init__(self, generator_net, discriminator_net, Lambda):
        super().__init__()
        self.generator_net = generator_net
        self.discriminator_net = discriminator_net
        self.Lambda = Lambda

    def forward(self, real_images, epsilon=None):
        if self.Lambda > 0:
            batch_size = real_images.size(0)
            device = real_images.device
            z_latent = sample_latent_vec((batch_size, self.generator_net.latent_dim), device=device)
            with torch.no_grad():
                x_tilde = self.generator_net(z_latent)
            if epsilon is None:
                epsilon = torch.rand((batch_size, 1, 1, 1), device=device)
            eps = epsilon.to(device=device, dtype=torch.float32).reshape(batch_size).contiguous()
            x_hat = ops.interp_images(real_images.to(torch.float32).contiguous(), x_tilde.contiguous(), eps)
            Gradient_penalty_loss = autograd_fns.gradient_penalty(self.discriminator_net, x_hat, self.Lambda)
        else:
            Gradient_penalty_loss = torch.tensor(0)
        return Gradient_penalty_loss


def similarity_loss(images_batch: torch.Tensor, Z_batch: torch.Tensor, Lambda: float = 1.0, data_parallel=None):
    """Optional anti-mode-collapse term (reference loss_functions.py:185-205; train.py:379-381 adds it to the generator
    loss when sim_loss_lambda > 0): Lambda / (B*(B-1)) * sum_ij (cos(z_i, z_j) - cos(x_i, x_j))^2.

    Runs in two kernels of libngan_b200.so (per-chunk Gram matrices, ordered reduction).  The Gram matrices couple the
    samples of a batch, so with data parallelism (torch.distributed initialised; `data_parallel` overrides) the rows of
    both tensors are all-gathered first and every rank returns the loss of the GLOBAL batch -- what a single process
    would compute.  As train.py calls it -- on the REAL images and a fresh latent draw -- nothing requires grad and the
    term only shifts the reported generator loss; inputs that do require grad take the differentiable tensor-algebra
    route below (same arithmetic as the reference, off the training path)."""
    import torch.distributed as dist
    if images_batch.requires_grad or Z_batch.requires_grad:
        batch_size = images_batch.size(0)
        images_mat = images_batch.view(batch_size, -1)
        Z_mat = Z_batch.view(batch_size, -1)
        images_mat = images_mat / images_mat.norm(2, dim=1, keepdim=True)
        Z_mat = Z_mat / Z_mat.norm(2, dim=1, keepdim=True)
        diff = torch.matmul(Z_mat, Z_mat.t()) - torch.matmul(images_mat, images_mat.t())
        return Lambda * torch.pow(diff, 2).sum() / (batch_size * (batch_size - 1))
    x = images_batch.detach().to(torch.float32).reshape(images_batch.size(0), -1).contiguous()
    z = Z_batch.detach().to(device=x.device, dtype=torch.float32).reshape(Z_batch.size(0), -1).contiguous()
    if data_parallel is None:
        data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if data_parallel:
        world = dist.get_world_size()
        xg = torch.empty((world * x.shape[0], x.shape[1]), dtype=x.dtype, device=x.device)
        zg = torch.empty((world * z.shape[0], z.shape[1]), dtype=z.dtype, device=z.device)
        dist.all_gather_into_tensor(xg, x)
        dist.all_gather_into_tensor(zg, z)
        x, z = xg, zg
    return ops.similarity_loss(x, z, float(Lambda))[0]
