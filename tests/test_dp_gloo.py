"""The N>1 path on CPU (gloo, world_size 2): row sharding of the global draws and the claim that averaging
per-rank gradients reproduces the single-process gradients exactly, checked with the
oracle as the per-rank compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuron_gan_b200 import dp
from neuron_gan_b200.utils import sample_latent_vec
from oracle import pggan_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, res, alpha, batch, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    r, w, _ = dp.init_from_env('gloo')
    assert (r, w) == (rank, world)
    arch = O.Arch()
    tr = O.Trainer(arch, seed=1, res=res, alpha=alpha)            # identical replicas on every rank
    x = dp.shard_rows(O.synthetic_images(batch, res))
    z1, z2, eps, z3 = dp.global_draws(sample_latent_vec, batch, 512)
    d_total, _, _, _ = tr.d_losses(x, z1, z2, eps)
    names = O.active_d_names(tr.n_layers, alpha, arch)
    grads = tr._grads(d_total, tr.dp, names)
    flat = torch.cat([grads[k].flatten() for k in names])
    dp.allreduce_mean_(flat)
    if rank == 0:
        out.put((names, flat.clone(), z1.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_average_equals_single_process():
    res, alpha, batch, world = 32, 0.5, 4, 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, res, alpha, batch, out)) for r in range(world)]
    for p in procs:
        p.start()
    names, flat, z1_rank0 = out.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single process, full batch, same global draws
    arch = O.Arch()
    tr = O.Trainer(arch, seed=1, res=res, alpha=alpha)
    x = O.synthetic_images(batch, res)
    z1, z2, eps, z3 = dp.global_draws(sample_latent_vec, batch, 512, rank=0, world=1)
    assert torch.equal(z1[:batch // world], z1_rank0)             # rank 0 consumed the same stream, bit-exactly
    d_total, _, _, _ = tr.d_losses(x, z1, z2, eps)
    grads = tr._grads(d_total, tr.dp, names)
    ref = torch.cat([grads[k].flatten() for k in names])
    assert torch.allclose(flat, ref, rtol=1e-4, atol=1e-7)


def test_shard_rows_rejects_ragged_batches():
    t = torch.arange(10).reshape(5, 2)
    with pytest.raises(ValueError):
        dp.shard_rows(t, 0, 2)
    assert torch.equal(dp.shard_rows(torch.arange(8).reshape(4, 2), 1, 2), torch.tensor([[4, 5], [6, 7]]))


@pytest.mark.parametrize('penalty', [True, False])
def test_global_draws_follow_the_reference_stream(penalty):
    """dp.global_draws consumes the CPU stream exactly as one process running the reference does: z, (z, eps), z --
    the middle pair only when the gradient penalty is on (loss_functions.py:159) -- and hands each rank its rows."""
    torch.manual_seed(5)
    shards = [dp.global_draws(sample_latent_vec, 8, 512, rank=0, world=2, penalty=penalty)]
    nxt = torch.rand(1).item()
    torch.manual_seed(5)
    shards.append(dp.global_draws(sample_latent_vec, 8, 512, rank=1, world=2, penalty=penalty))
    torch.manual_seed(5)
    z1 = sample_latent_vec((8, 512))
    if penalty:
        z2, eps = sample_latent_vec((8, 512)), torch.rand((8, 1, 1, 1))
    else:
        z2, eps = torch.zeros(8, 512), torch.zeros(8, 1, 1, 1)
    z3 = sample_latent_vec((8, 512))
    assert torch.rand(1).item() == nxt
    for i, full in enumerate((z1, z2, eps, z3)):
        assert torch.equal(torch.cat([shards[0][i], shards[1][i]]), full)
