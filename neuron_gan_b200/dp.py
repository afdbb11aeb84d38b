"""Data-parallel plumbing (new functionality: the reference is single-device).

One process per GPU; rank r owns rows [r*B/P, (r+1)*B/P) of the global image batch and of every global RNG draw
(z, z, eps, z), so any GPU count consumes the single-process random stream bit-exactly.
Gradients are averaged with one all-reduce per flat buffer per optimiser step (NCCL over NVLink on GPUs; the
same code runs over gloo on CPU tensors in the tests).  Exact for this loss: every term is a batch mean of
per-sample quantities and no layer couples samples."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Returns
    (rank, world, local_rank).  No-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def shard_rows(t, rank=None, world=None):
    """This rank's contiguous slice of dim 0 of a globally drawn tensor."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    n = t.shape[0]
    if n % world:
        raise ValueError(f'global batch {n} is not divisible by the world size {world}')
    per = n // world
    return t[rank * per:(rank + 1) * per]


def global_draws(sample_latent_vec, global_batch, latent_dim, rank=None, world=None, penalty=True):
    """Draw (z, z, eps, z) for the GLOBAL batch on the CPU generator in the reference's order
    (loss_functions.py:25, 166, 170, 63) and return this rank's rows.  penalty=False (grad_pen_lambda == 0): the
    reference's D_grad_pen_loss then draws nothing (loss_functions.py:159), so neither does this -- its z and eps
    come back as zeros."""
    z1 = sample_latent_vec((global_batch, latent_dim))
    if penalty:
        z2 = sample_latent_vec((global_batch, latent_dim))
        eps = torch.rand((global_batch, 1, 1, 1))
    else:
        z2, eps = torch.zeros((global_batch, latent_dim)), torch.zeros((global_batch, 1, 1, 1))
    z3 = sample_latent_vec((global_batch, latent_dim))
    return tuple(shard_rows(t, rank, world).contiguous() for t in (z1, z2, eps, z3))


def allreduce_mean_(flat):
    """In-place mean over ranks of a flat gradient buffer."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if flat.is_cuda:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        else:                                   # gloo has no AVG
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(dist.get_world_size())
    return flat
