"""TrainStep: one iteration of the reference's pggan_train inner loop (reference train.py:350-394, n_critic = 1)
driven straight through the engine -- the call a training launcher makes once per batch.

    D step:  D.zero_grad(); D_W_loss + D_grad_pen_loss; backward; Adam(D)        (train.py:356-366)
    G step:  G.zero_grad(); G_W_loss; backward; Adam(G)                         (train.py:375-385)

It produces the same numbers as calling the loss modules of loss_functions.py + loss.backward() + FusedAdam
(tests/test_step_gpu.py checks that), minus the per-call autograd bookkeeping and host syncs: gradients
accumulate in one flat fp32 buffer per network (`p.grad` are views of it), the six `.item()` calls of
train.py:389-394 become one packed device tensor, and the NaN guards read that tensor one step late.

Data parallel: with torch.distributed initialised, every rank runs the same step on its
shard of the batch and the flat gradient buffers are all-reduced (mean) over NCCL before each Adam step, which
is exact for this loss (batch means of per-sample terms, no cross-sample layers).
"""
import math

import torch
import torch.distributed as dist

from . import engine, ops
from .optim import FusedAdam
from .utils import sample_latent_vec

F32 = torch.float32


class TrainStep:
    def __init__(self, generator_net, discriminator_net, learning_rate=1e-4, beta1=0.5, grad_pen_lambda=10.0,
                 drift_epsilon=1e-3, data_parallel=None):
        self.G, self.D = generator_net, discriminator_net
        self.lam, self.drift = float(grad_pen_lambda), float(drift_epsilon)
        self.opt_g = FusedAdam(self.G.parameters(), lr=learning_rate, betas=(beta1, 0.999))
        self.opt_d = FusedAdam(self.D.parameters(), lr=learning_rate, betas=(beta1, 0.999))
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.dp = data_parallel
        self._bound = {}
        self.kernel_launches = 0
        self._pending_stats = None

    # -- flat gradient buffers ---------------------------------------------------------------------------
    def _bind(self, net):
        active = net.active_parameters()
        key = tuple(id(p) for p in active)
        ent = self._bound.get(id(net))
        if ent is None or ent['key'] != key:
            n = sum(p.numel() for p in active)
            flat = torch.zeros(n, dtype=F32, device=active[0].device)
            sink, off = {}, 0
            for p in net.parameters():
                p.grad = None
            # the generator's Linear weight goes last: it is produced last in backward and is its own bucket
            for p in sorted(active, key=lambda q: q.numel() > (1 << 22)):
                v = flat[off:off + p.numel()].view(p.shape)
                p.grad = v
                sink[id(p)] = v
                off += p.numel()
            ent = {'key': key, 'flat': flat, 'sink': sink}
            self._bound[id(net)] = ent
        return ent['flat'], ent['sink']

    def _allreduce(self, flat):
        if self.dp:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)

    # -- RNG draws in the reference's order: z (D_W_loss) -> z (grad pen) -> eps -> z (G_W_loss) --------------
    def draw(self, batch, device):
        z1 = sample_latent_vec((batch, self.G.latent_dim))
        z2 = sample_latent_vec((batch, self.G.latent_dim))
        eps = torch.rand((batch, 1, 1, 1))          # CPU generator for reproducibility
        z3 = sample_latent_vec((batch, self.G.latent_dim))
        return tuple(t.to(device, non_blocking=True) for t in (z1, z2, eps, z3))

    @torch.no_grad()
    def __call__(self, images, draws=None):
        """images: [B, 1, R, R] fp32 on the GPU (this rank's shard).  Returns a device tensor
        [D_loss, score_real, score_fake, G_loss, D_grad_pen] (D_loss includes the penalty, train.py:362)."""
        G, D = self.G, self.D
        dev = images.device
        B = images.shape[0]
        x = images.reshape(B, images.shape[-2], images.shape[-1]).to(F32).contiguous()
        z1, z2, eps, z3 = draws if draws is not None else self.draw(B, dev)
        eps = eps.reshape(B).to(F32).contiguous()

        # ---------------- critic step (train.py:356-366)
        flat_d, sink_d = self._bind(D)
        flat_d.zero_()
        fake, _ = engine.g_forward(G, z1, save=False)
        scores, ctx = engine.d_forward(D, torch.cat([x, fake]), save=True)
        gout = torch.empty(2 * B, dtype=F32, device=dev)
        out3, _, _ = ops.wloss_into(scores[:B], scores[B:], self.drift, gout[:B], gout[B:])
        engine.d_backward(D, ctx, gout, sink_d)
        del ctx
        x_tilde, _ = engine.g_forward(G, z2, save=False)
        x_hat = ops.interp_images(x, x_tilde, eps)
        pen, _, _ = engine.d_grad_penalty(D, x_hat, self.lam, sink_d)
        self._allreduce(flat_d)
        self.opt_d.step()

        # ---------------- generator step (train.py:375-385)
        flat_g, sink_g = self._bind(G)
        flat_g.zero_()
        fake, gctx = engine.g_forward(G, z3, save=True)
        s_fake, dctx = engine.d_forward(D, fake, save=True)
        out1, g_fake = ops.gloss(s_fake)
        g_xp = engine.d_backward(D, dctx, g_fake, None, want_gxp=True)
        gx = ops.unpool_image(g_xp, 0.25) if dctx.pooled else g_xp
        del dctx
        engine.g_backward(G, gctx, gx, sink_g)
        del gctx
        self._allreduce(flat_g)
        self.opt_g.step()

        stats = torch.cat([out3, out1, pen])
        stats[0] += stats[4]
        return stats

    @staticmethod
    def stats_dict(stats_host):
        v = [float(t) for t in stats_host]
        return {'D_loss': v[0], 'score_real': v[1], 'score_fake': v[2], 'G_loss': v[3], 'D_grad_pen': v[4]}

    @staticmethod
    def check_nan(stats_host):
        """The reference's NaN guards (loss_functions.py:35-41, 70-72), applied to the packed statistics."""
        if math.isnan(stats_host[1]):
            raise ValueError('Real loss is nan.')
        if math.isnan(stats_host[2]):
            raise ValueError('Fake loss is nan.')
        if math.isnan(stats_host[3]):
            raise ValueError('Generator loss is nan.')


def build_networks(res=16, alpha=1.0, seed=1, device='cuda', gen_features=None, dis_features=None, image_size=512):
    """torch.manual_seed(seed) -> Generator_PG -> Discriminator_PG -> set_resolution (train.py:114, 172, 184)."""
    from .models import Discriminator_PG, Generator_PG
    gen_features = gen_features or [128, 64, 32, 32, 16, 16]
    dis_features = dis_features or [16, 16, 32, 32, 64, 128]
    size_init = image_size // 2 ** (len(gen_features) - 1)
    torch.manual_seed(seed)
    G = Generator_PG(list(gen_features), image_size_init=size_init)
    D = Discriminator_PG(list(dis_features), image_size_init=size_init)
    if res != size_init:
        G.set_resolution(res, alpha)
        D.set_resolution(res, alpha)
    return G.to(device), D.to(device)


def smoke_check(device, res=16, batch=4, tol=1e-2):
    """One tiny iteration on the GPU, compared with the CPU oracle on the same seeds (used by
    __graft_entry__.smoke; the oracle is only the checker here)."""
    from oracle import pggan_oracle as O
    arch = O.Arch()
    tr = O.Trainer(arch, seed=1, res=res, alpha=1.0)
    rng = torch.get_rng_state()
    x = O.synthetic_images(batch, res)
    ref = tr.iteration(x)
    G, D = build_networks(res, 1.0, seed=1, device=device)
    step = TrainStep(G, D)
    torch.set_rng_state(rng)
    stats = step(x.to(device)).cpu()
    got = TrainStep.stats_dict(stats)
    for k, v in ref.items():
        if abs(got[k] - v) > tol * max(1.0, abs(v)):
            raise AssertionError(f'smoke: {k} = {got[k]} differs from the oracle {v}')
    print('smoke ok:', got)
    return got
