"""Data-parallel equivalence on the CUDA path (SURVEY.md section 4 tier 6): N GPUs x B/N images against 1 GPU x B
images on the same global images and draws, through TrainStep.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dp_equivalence.py [res] [alpha] [global batch] [steps]

Every rank draws the GLOBAL batch's latents / epsilon on its identically seeded CPU generator and keeps its rows; the
images are sliced the same way.  Rank 0 then repeats the run alone on the global batch.  Checked and printed:
  * the replicas are bit-identical after the DP run (same averaged gradients, same Adam);
  * DP statistics (mean over ranks of the per-shard statistics) == single-process statistics;
  * every parameter after `steps` iterations: max / mean |difference| between the DP and the single-process run
    (gradients of shards are summed in a different order than one big batch, so this is close, not bitwise)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuron_gan_b200 import dp  # noqa: E402
from neuron_gan_b200.train_step import TrainStep, build_networks  # noqa: E402
from neuron_gan_b200.utils import sample_latent_vec  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 64
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
GB = int(sys.argv[3]) if len(sys.argv) > 3 else 16
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4

rank, world, local = dp.init_from_env('nccl')
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
gen = torch.Generator().manual_seed(11)
images = [torch.rand(GB, 1, res, res, generator=gen) * 2 - 1 for _ in range(steps)]


def run(data_parallel):
    G, D = build_networks(res, alpha, seed=1, device=dev)
    step = TrainStep(G, D, data_parallel=data_parallel)
    torch.manual_seed(77)                     # the CPU draw stream, identical on every rank
    out = []
    for x in images:
        if data_parallel:
            xs = dp.shard_rows(x, rank, world).contiguous().to(dev)
            out.append(step(xs).cpu())        # TrainStep draws the global batch and keeps this rank's rows
        else:
            draws = tuple(t.to(dev) for t in dp.global_draws(sample_latent_vec, GB, G.latent_dim, 0, 1))
            out.append(step(x.to(dev), draws).cpu())
    torch.cuda.synchronize()
    return G, D, torch.stack(out), step


G, D, stats_dp, step_dp = run(True)
# statistics: mean over ranks of the shard statistics
s = stats_dp.to(dev)
dist.all_reduce(s, op=dist.ReduceOp.AVG)
stats_dp = s.cpu()
digest = torch.stack([p.detach().double().sum() for net in (G, D) for p in net.parameters()] +
                     [p.detach().double().abs().sum() for net in (G, D) for p in net.parameters()])
digests = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(digests, digest)
identical = all(torch.equal(d, digests[0]) for d in digests)
dist.barrier()
if rank == 0:
    G1, D1, stats_1, _ = run(False)
    print(f'world={world} res={res} alpha={alpha} global batch={GB} steps={steps} last_run={step_dp.last_run}')
    print('replicas bit-identical after the DP run:', identical)
    print('stats DP     :', [round(v, 6) for v in stats_dp[-1].tolist()])
    print('stats 1 GPU  :', [round(v, 6) for v in stats_1[-1].tolist()])
    dmax = (stats_dp - stats_1).abs().max().item()
    print(f'max |stat difference| over {steps} iterations: {dmax:.3e}')
    worst, mean = 0.0, 0.0
    n = 0
    for (k, a), (_, b) in list(zip(G.state_dict().items(), G1.state_dict().items())) + \
            list(zip(D.state_dict().items(), D1.state_dict().items())):
        d = (a.float() - b.float()).abs()
        worst = max(worst, d.max().item())
        mean += d.sum().item()
        n += d.numel()
    print(f'parameters after {steps} iterations: max |diff| {worst:.3e}, mean |diff| {mean / n:.3e} (lr = 1e-4)')
    ok = identical and dmax < 5e-3 and worst <= steps * 2.1e-4 and mean / n < 1e-5
    print('DP_EQUIVALENCE_OK' if ok else 'DP_EQUIVALENCE_FAILED')
dist.barrier()
dist.destroy_process_group()
