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
from types import SimpleNamespace

import torch
import torch.distributed as dist

from . import _lib, engine, ops
from .optim import FusedAdam
from .utils import PinnedRing, sample_latent_vec

F32 = torch.float32


class TrainStep:
    """use_graph: None = capture the iteration into a CUDA graph as soon as the same configuration (batch,
    resolution, alpha, active parameters) has been seen twice in a row, and replay it from then on; False = always
    launch kernel by kernel.  Captured or not, the kernels and their order are identical."""

    def __init__(self, generator_net, discriminator_net, learning_rate=1e-4, beta1=0.5, grad_pen_lambda=10.0,
                 drift_epsilon=1e-3, data_parallel=None, use_graph=None, n_critic=1):
        self.G, self.D = generator_net, discriminator_net
        self.lam, self.drift = float(grad_pen_lambda), float(drift_epsilon)
        self.opt_g = FusedAdam(self.G.parameters(), lr=learning_rate, betas=(beta1, 0.999), capturable=True)
        self.opt_d = FusedAdam(self.D.parameters(), lr=learning_rate, betas=(beta1, 0.999), capturable=True)
        if data_parallel is None:
            data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.dp = data_parallel
        self.use_graph = use_graph
        self.n_critic = int(n_critic)   # critic steps per generator step (train.py:356; > 1 runs kernel by kernel)
        self.segment_graphs = False     # capture three graphs even on one GPU (lets a caller time the three parts)
        self.segment_events = None      # when a list: 4 CUDA events per replayed iteration are appended (start, D, G, end)
        self.global_draws = True        # data parallel: draw the global batch on every rank and keep this rank's rows
        # The generator's Linear weight is 98 % of its parameters and its gradient a rank-B product ga0^T z.
        # factor_linear: keep the gradient as its two factors after the backward pass.  With data parallelism the
        # ranks then all-gather 1 MB of factors instead of all-reducing the 67 MB product, and form the global
        # gradient locally on tensor cores (ops.linear_wgrad_factored).  fuse_linear_adam: form it inside the Adam pass
        # and never write it (ops.adam_linear_factored; lin.weight.grad then stays unmaterialised unless
        # materialize_linear_grad) -- measured 2 % slower than the separate kernels on one GPU, so off by default.
        self.factor_linear = True
        self.fuse_linear_adam = False
        self.materialize_linear_grad = False
        self.merge_g_forward = None     # run the iteration's three generator passes (z1, z2 detached; z3 with saved
                                        # activations) as ONE batch: one chain of launches instead of two (measured,
                                        # graph replay: 16x16 0.636 -> 0.617 ms, 128x128 2.907 -> 2.869, 512x512 3.537
                                        # -> 3.494 although the activations of 2B samples are then saved for nothing).
                                        # None = always.  (With data parallelism a separate third pass could hide the
                                        # critic's 2 MB gradient all-reduce, but measured at 2 GPUs the merged pass wins:
                                        # set False to get the overlapped arrangement.)
        self.fork_chains = True         # run the two independent halves of the critic step on two streams
        self._chain_stream = None
        self._comm_stream = None        # data parallel: the critic's gradient all-reduce overlaps the generator forward
        self._bound = {}
        self._graphs = {}               # configuration key -> captured iteration, least recently used first
        self.max_graphs = 4             # during a fade-in alpha (part of the key) changes every epoch: old graphs go
        self._rings = {}
        self._last_key = None
        self._versions_seen = None
        self._alpha_keep = []           # pinned sources of the last alpha uploads (alive until the copies have run)
        self.last_run = None            # 'eager' | 'replay': how the last iteration was executed
        self.launches_per_step = 0      # kernels of libngan_b200.so launched (or replayed) by the last iteration
        self.inputs_loaded = None       # CUDA event: this iteration's host-to-device copies (images, draws, Adam scalars)
                                        # are done -- a loader may start its next big copy behind it (DevicePrefetcher)

    # -- flat gradient buffers ---------------------------------------------------------------------------
    def _bind(self, net, n_slots=1):
        """One flat fp32 gradient buffer per network (`p.grad` are views of it -> one all-reduce, one Adam launch).
        Every gradient kernel OVERWRITES its parameter's tensor without atomics (bit-reproducible).  The critic's
        weights receive three contributions per step that run concurrently -- the Wasserstein backward and the two
        sweeps of the penalty's double backward -- so the critic gets three slot buffers (one sink each) that
        ops.sum_slots adds in a fixed order into the flat buffer; (parameter, slot) pairs no kernel writes stay zero
        from allocation.  Nothing is zeroed per step (the reference's zero_grad(), train.py:357, 375)."""
        active = net.active_parameters()
        key = tuple(id(p) for p in active)
        ent = self._bound.get(id(net))
        if ent is None or ent['key'] != key:
            n = sum(p.numel() for p in active)
            dev = active[0].device
            flat = torch.zeros(n, dtype=F32, device=dev)
            slots = torch.zeros((n_slots, n), dtype=F32, device=dev) if n_slots > 1 else flat.view(1, n)
            sinks, off = [engine.Sink(accumulate=False) for _ in range(n_slots)], 0
            for p in net.parameters():
                p.grad = None
            # the generator's Linear weight goes last: it is produced last in backward and is its own bucket
            for p in sorted(active, key=lambda q: q.numel() > (1 << 22)):
                p.grad = flat[off:off + p.numel()].view(p.shape)
                for k in range(n_slots):
                    sinks[k][id(p)] = slots[k, off:off + p.numel()].view(p.shape)
                off += p.numel()
            big = sum(p.numel() for p in active if p.numel() > (1 << 22))
            ent = {'key': key, 'flat': flat, 'sinks': sinks, 'slots': slots if n_slots > 1 else None,
                   'small': flat[:n - big] if big else flat}
            self._bound[id(net)] = ent
        return ent['flat'], ent['sinks']

    def _allreduce(self, flat):
        if self.dp:
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)

    def _exchange_g(self, buf):
        """Data-parallel exchange after the generator's backward pass, as ONE collective: every rank packs
        [small gradients | latents z | Linear-gradient factor ga0] into one buffer, the buffers are all-gathered, and
        each rank (a) adds the ranks' small gradients in rank order (ops.sum_slots: every replica computes the same
        bits, whatever algorithm NCCL picks) and (b) hands the gathered factors to ops.adam_linear_factored as
        per-rank segments.  1.2 MB + 32 KB + 1 MB per rank instead of an all-reduce of 68 MB."""
        if not self.dp:
            return
        ent = self._bound[id(self.G)]
        if not self.factor_linear:
            dist.all_reduce(ent['flat'], op=dist.ReduceOp.AVG)
            return
        pk = buf.pack
        pk.send_small.copy_(ent['small'], non_blocking=True)
        pk.send_z.copy_(buf.out.lin.z.reshape(-1), non_blocking=True)
        pk.send_ga.copy_(buf.out.lin.ga0.reshape(-1), non_blocking=True)
        dist.all_gather_into_tensor(pk.recv.view(-1), pk.send)
        ops.sum_slots(pk.recv_small, ent['small'], scale=1.0 / pk.world)

    def _make_pack(self, B, dev):
        """Send / receive buffers of _exchange_g (per step-buffer set: the factor sizes depend on the batch)."""
        world = dist.get_world_size()
        n_small = self._bind(self.G)[0].numel() - sum(p.numel() for p in self.G.active_parameters()
                                                      if p.numel() > (1 << 22))
        L = self.G.latent_dim
        n_ga = B * self.G.N_features_per_layer[0] * self.G.image_size_init ** 2
        pad = lambda nbytes: (nbytes + 15) // 16 * 16
        o_z = pad(n_small * 4)
        o_ga = o_z + pad(B * L * 4)
        per = o_ga + pad(n_ga * 2)
        send = torch.empty(per, dtype=torch.uint8, device=dev)
        recv = torch.empty((world, per), dtype=torch.uint8, device=dev)
        view = lambda t, o, n, dt: t[..., o:o + n * (4 if dt is F32 else 2)].view(dt)
        return SimpleNamespace(world=world, per=per, send=send, recv=recv, n_small=n_small,
                               send_small=view(send, 0, n_small, F32), send_z=view(send, o_z, B * L, F32),
                               send_ga=view(send, o_ga, n_ga, torch.bfloat16),
                               recv_small=view(recv, 0, n_small, F32),              # [world, n_small], row stride per / 4
                               recv_z=view(recv, o_z, B * L, F32), recv_ga=view(recv, o_ga, n_ga, torch.bfloat16))

    # -- RNG draws in the reference's order: z (D_W_loss) -> z (grad pen) -> eps -> z (G_W_loss) --------------
    def draw_host(self, batch):
        """This rank's rows of the four draws.  With data parallelism every rank draws the GLOBAL batch on its
        (identically seeded) CPU generator in the reference's order and keeps rows [rank*b, (rank+1)*b), so any GPU
        count consumes the single-process random stream bit-exactly (dp.global_draws)."""
        if self.dp and dist.is_initialized() and self.global_draws:
            from .dp import global_draws
            return global_draws(sample_latent_vec, batch * dist.get_world_size(), self.G.latent_dim,
                                penalty=self.lam > 0)
        z1 = sample_latent_vec((batch, self.G.latent_dim))
        if self.lam > 0:
            z2 = sample_latent_vec((batch, self.G.latent_dim))
            eps = torch.rand((batch, 1, 1, 1))      # CPU generator for reproducibility
        else:
            # grad_pen_lambda == 0: the reference's D_grad_pen_loss returns 0 WITHOUT drawing (loss_functions.py:159);
            # the penalty chain still runs here with coefficient 0 (zero loss, zero gradients) on placeholder inputs
            z2, eps = torch.zeros((batch, self.G.latent_dim)), torch.zeros((batch, 1, 1, 1))
        z3 = sample_latent_vec((batch, self.G.latent_dim))
        return z1, z2, eps, z3

    def draw_critic_host(self, batch):
        """(z, z, eps) of one critic round (loss_functions.py:25, 166, 170) for `batch` rows; no penalty draws when
        grad_pen_lambda == 0 (loss_functions.py:159)."""
        z1 = sample_latent_vec((batch, self.G.latent_dim))
        if self.lam > 0:
            return z1, sample_latent_vec((batch, self.G.latent_dim)), torch.rand((batch, 1, 1, 1))
        return z1, torch.zeros((batch, self.G.latent_dim)), torch.zeros((batch, 1, 1, 1))

    def draw(self, batch, device):
        """The four draws of one iteration as device tensors (for callers that keep draws resident)."""
        return tuple(t.to(device) for t in self.draw_host(batch))

    # -- the iteration as three kernel sequences; gradients are complete at the end of seg_d and seg_g -----------
    def _buffers(self, B, R, dev):
        """Step inputs: imgs = [x | G(z1) | G(z2)] (the critic batch [real; fake] and x_tilde are views of it),
        z12 = [z1; z2] (both detached generator passes of the critic step run as one batch)."""
        L = self.G.latent_dim
        z_all = torch.empty((3 * B, L), dtype=F32, device=dev)          # [z1; z2; z3]
        buf = SimpleNamespace(imgs=torch.empty((4 * B, R, R), dtype=F32, device=dev),   # [x | G(z1) | G(z2) | G(z3)]
                              z_all=z_all, z12=z_all[:2 * B], z3=z_all[2 * B:],
                              eps=torch.empty((B,), dtype=F32, device=dev), B=B, R=R, out=SimpleNamespace())
        if self.dp and self.factor_linear:
            buf.pack = self._make_pack(B, dev)
        return buf

    def _load(self, buf, x, z1, z2, eps, z3):
        """Copy the iteration's inputs into the step buffers.  Device and pinned-host sources are copied directly
        (asynchronously); pageable host tensors -- the CPU draws -- go through a ring of pinned staging buffers."""
        B, L = buf.B, self.G.latent_dim
        buf.imgs[:B].copy_(x.reshape(B, x.shape[-2], x.shape[-1]), non_blocking=True)
        pairs = ((buf.z12[:B], z1), (buf.z12[B:], z2), (buf.eps, eps.reshape(B)), (buf.z3, z3))
        if all(src.is_cuda or src.is_pinned() for _, src in pairs):
            for dst, src in pairs:
                dst.copy_(src, non_blocking=True)
            return
        ring = self._rings.get(B)
        if ring is None:
            ring = self._rings[B] = PinnedRing([(B, L), (B, L), (B,), (B, L)])
        k, stage = ring.acquire()
        for (dst, src), st in zip(pairs, stage):
            st.copy_(src)
            dst.copy_(st, non_blocking=True)
        ring.release(k)

    def _seg_d(self, buf):
        """critic step up to complete gradients (train.py:356-365).  After the shared generator pass the
        Wasserstein part ([real; fake] forward + backward) and the gradient-penalty part (x_hat forward,
        first-order backward, double backward) are independent chains that write separate gradient slots (summed in
        order afterwards): they run on two streams, so the narrow low-resolution kernels of one fill the SMs the other
        leaves idle."""
        G, D, B = self.G, self.D, buf.B
        flat_d, (sink_w, sink_s1, sink_s2) = self._bind(D, 3)
        if self._merged(buf):
            _, gctx = engine.g_forward(G, buf.z_all, save=True, img_out=buf.imgs[B:])
            buf.out.fake, buf.out.gctx = buf.imgs[3 * B:], engine.g_ctx_slice(gctx, 2 * B, 3 * B)
            del gctx
        else:
            engine.g_forward(G, buf.z12, save=False, img_out=buf.imgs[B:3 * B])
        main = torch.cuda.current_stream()
        fork = self.fork_chains
        if fork:
            if self._chain_stream is None:
                self._chain_stream = torch.cuda.Stream()
            engine.prepare_weights(D)
            self._chain_stream.wait_stream(main)
        with torch.cuda.stream(self._chain_stream if fork else main):
            x_hat = ops.interp_images(buf.imgs[:B], buf.imgs[2 * B:3 * B], buf.eps)
            buf.out.pen, _, _ = engine.d_grad_penalty(D, x_hat, self.lam, (sink_s1, sink_s2))
            del x_hat
        scores, ctx = engine.d_forward(D, buf.imgs[:2 * B], save=True)
        gout = torch.empty(2 * B, dtype=F32, device=scores.device)
        buf.out.out3, _, _ = ops.wloss_into(scores[:B], scores[B:], self.drift, gout[:B], gout[B:])
        engine.d_backward(D, ctx, gout, sink_w)
        del ctx
        if fork:
            main.wait_stream(self._chain_stream)
        engine.side_join()
        ops.sum_slots(self._bound[id(D)]['slots'], flat_d)
        return flat_d

    def _seg_g1(self, buf):
        """The generator's forward pass of the generator step (train.py:376).  It depends on neither the critic's
        gradients nor its update, so with data parallelism it runs while the critic's gradients are all-reduced."""
        if not self._merged(buf):
            buf.out.fake, buf.out.gctx = engine.g_forward(self.G, buf.z3, save=True)

    def _merged(self, buf):
        if self.n_critic != 1:
            return False
        return True if self.merge_g_forward is None else bool(self.merge_g_forward)

    def _seg_g(self, buf, adam_d=True):
        self._seg_g1(buf)
        return self._seg_g2(buf, adam_d)

    def _seg_g2(self, buf, adam_d=True):
        """Adam(D) (train.py:366), then the rest of the generator step up to complete gradients (train.py:375-384)"""
        G, D = self.G, self.D
        if adam_d:
            self.opt_d.launch()
        flat_g, (sink_g,) = self._bind(G)
        fake, gctx = buf.out.fake, buf.out.gctx
        buf.out.fake = buf.out.gctx = None
        s_fake, dctx = engine.d_forward(D, fake, save=True)
        buf.out.out1, g_fake = ops.gloss(s_fake)
        g_xp = engine.d_backward(D, dctx, g_fake, None, want_gxp=True)
        gx = ops.unpool_image(g_xp, 0.25) if dctx.pooled else g_xp
        del dctx
        buf.out.lin = SimpleNamespace() if self.factor_linear else None
        engine.g_backward(G, gctx, gx, sink_g, linear_overwrite=True, linear_factors=buf.out.lin)
        del gctx
        engine.side_join()
        return flat_g

    def _seg_end(self, buf):
        """Adam(G) (train.py:385) and the packed statistics (train.py:362, 389-394)"""
        factored = None
        if self.factor_linear:
            lin, f = self.G.layers[0], buf.out.lin
            K, C, S = lin._ngan_dims
            world = dist.get_world_size() if self.dp else 1
            dw = self._bound[id(self.G)]['sinks'][0][id(lin.weight)]
            d = dict(K=K, C=C, S=S, gscale=f.scale / world, b_per_seg=buf.B, n_seg=world)
            if self.dp:
                d.update(ga=buf.pack.recv_ga, z=buf.pack.recv_z, ga_seg_stride=buf.pack.per, z_seg_stride=buf.pack.per)
            else:
                d.update(ga=f.ga0, z=f.z)
            if self.fuse_linear_adam:
                d['g_out'] = dw if self.materialize_linear_grad else None
                factored = {id(lin.weight): d}
            else:
                ops.linear_wgrad_factored(d['ga'], d['z'], K, C, S, d['gscale'], dw, buf.B, world,
                                          d.get('ga_seg_stride', 0), d.get('z_seg_stride', 0))
        self.opt_g.launch(factored=factored)
        stats = torch.empty(5, dtype=F32, device=buf.z3.device)
        ops.pack_stats(buf.out.out3, buf.out.out1, buf.out.pen, stats)
        self._last_critic_out = (buf.out.out3, buf.out.pen)
        return stats

    def _run_eager(self, buf):
        n0 = _lib.launch_count
        stats = self._run_eager_body(buf)
        self.launches_per_step = _lib.launch_count - n0
        return stats

    def _run_eager_body(self, buf):
        self._bind(self.D, 3)       # p.grad views must exist before advance(): it skips parameters without grad
        self._bind(self.G)
        self.opt_d.advance()
        self.opt_g.advance()
        self._allreduce(self._seg_d(buf))
        self._seg_g(buf)
        self._exchange_g(buf)
        return self._seg_end(buf)

    def _run_multi_critic(self, images, draws, dev):
        """n_critic > 1 (train.py:356-366): every critic step takes the same images and fresh draws (z, z, eps);
        the statistics are those of the last critic step, as in the reference.  draws: None, or a list of n_critic
        (z1, z2, eps) tuples followed by one z3 (n_critic = 0: [z3], or [(z1, z2, eps), z3] for the monitoring-only
        evaluation of the critic losses)."""
        B, R = images.shape[0], images.shape[-1]
        buf = self._buffers(B, R, dev)
        self._bind(self.D, 3)
        self._bind(self.G)
        for j in range(self.n_critic):
            z1, z2, eps = self.draw_critic_host(B) if draws is None else draws[j]
            self._load(buf, images, z1, z2, eps, z1)          # (z3 slot: overwritten below)
            self.opt_d.advance()
            self._allreduce(self._seg_d(buf))
            self.opt_d.launch()
            self._last_critic_out = (buf.out.out3, buf.out.pen)
        if self.n_critic == 0:
            # adapt_critic can ask for no critic step (train.py:336-340 with N_min = 0).  The reference then still
            # EVALUATES the critic losses for its statistics (train.py:369-374) -- consuming the draws z, z, eps --
            # without backward or optimiser step.  Here the critic segment runs as usual (its gradients stay in the
            # critic's buffers and are never applied: no all-reduce, no Adam launch, no step-count advance).
            z1, z2, eps = self.draw_critic_host(B) if draws is None or len(draws) < 2 else draws[0]
            self._load(buf, images, z1, z2, eps, z1)
            self._seg_d(buf)
        z3 = sample_latent_vec((B, self.G.latent_dim)) if draws is None else draws[-1]
        buf.z3.copy_(z3.to(dev) if not z3.is_cuda else z3)
        self.opt_g.advance()
        self._seg_g(buf, adam_d=False)
        self._exchange_g(buf)
        self._versions_seen = None
        return self._seg_end(buf)

    def _sync_alpha(self, dev):
        """Keep [alpha, 1 - alpha] of each network in device memory (engine._alpha_terms): the fade-in coefficient
        advances every epoch (train.py:318-321) and must not force a new graph each time."""
        for net in (self.G, self.D):
            a = net.alpha_value()
            if getattr(net, '_ngan_alpha_dev', None) is None or net._ngan_alpha_dev.device != dev:
                net._ngan_alpha_dev = torch.zeros(2, dtype=F32, device=dev)
                net._ngan_alpha_dev_value = None
            if net._ngan_alpha_dev_value != a:
                src = torch.tensor([a, 1.0 - a], dtype=torch.float64).to(F32).pin_memory()
                net._ngan_alpha_dev.copy_(src, non_blocking=True)
                net._ngan_alpha_dev_value = a
                self._alpha_keep = self._alpha_keep[-3:] + [src]

    # -- CUDA-graph capture ----------------------------------------------------------------------------------
    def _versions(self):
        # (_ngan_ext: engine.invalidate(), for writes through p.data that do not bump p._version)
        return tuple((p._version, getattr(p, '_ngan_ext', 0)) for net in (self.G, self.D) for p in net.parameters())

    def _config_key(self, B, R):
        # alpha itself is NOT part of the key: the kernels read it from device memory (_sync_alpha), so one graph
        # serves every epoch of a fade-in; only whether a fade is in progress changes the kernel sequence
        return (B, R, self.G.alpha_value() < 1, self.D.alpha_value() < 1, self.G.N_layers, self.D.N_layers, self.dp,
                self.lam, self.drift, self.factor_linear, self.fuse_linear_adam, self.materialize_linear_grad,
                self.merge_g_forward,
                tuple(id(p) for p in self.G.active_parameters()), tuple(id(p) for p in self.D.active_parameters()))

    def _capture(self, key, B, R, dev):
        buf = self._buffers(B, R, dev)
        ent = SimpleNamespace(buf=buf, graphs=[], flats=[], stats=None, launches=0)
        torch.cuda.synchronize()
        n0 = _lib.launch_count
        def whole(b):      # single GPU: the whole iteration is one graph
            self._seg_d(b)
            self._seg_g(b)
            return self._seg_end(b)

        def capture(segs):
            pool, graphs, outs = None, [], []
            for seg in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    outs.append(seg(buf))
                pool = g.pool()
                graphs.append(g)
            return graphs, outs

        # With data parallelism: four graphs, the collectives issued by the host in between.  (Capturing the NCCL
        # calls into one graph was measured: same iteration time at 2 GPUs, and the process then hung in
        # destroy_process_group -- not worth it.)
        parts = [self._seg_d, self._seg_g1, self._seg_g2, self._seg_end]
        if self._merged(buf):
            parts.pop(1)                # the generator's forward pass already ran with the critic step's
        if self.dp or self.segment_graphs:
            ent.graphs, ent.flats = capture(parts)
        else:
            ent.graphs, ent.flats = capture([whole])
        ent.stats = ent.flats[-1]
        ent.launches = _lib.launch_count - n0
        _lib.launch_count = n0          # captured, not launched
        self._graphs[key] = ent
        while len(self._graphs) > max(1, self.max_graphs):       # dicts keep insertion order: first = least recent
            self._graphs.pop(next(iter(self._graphs)))
        return ent

    @torch.no_grad()
    def __call__(self, images, draws=None):
        """images: [B, 1, R, R] fp32 (this rank's shard; GPU tensor, or pinned host tensor).  Returns a device tensor
        [D_loss, score_real, score_fake, G_loss, D_grad_pen] (D_loss includes the penalty, train.py:362)."""
        dev = next(self.G.parameters()).device
        if self.n_critic != 1:
            self._sync_alpha(dev)
            self.last_run = 'eager'
            return self._run_multi_critic(images, draws, dev)
        self._sync_alpha(dev)
        B, R = images.shape[0], images.shape[-1]
        z1, z2, eps, z3 = draws if draws is not None else self.draw_host(B)
        key = self._config_key(B, R)
        ent = self._graphs.get(key)
        versions = self._versions()
        if self.use_graph is False or ent is None or versions != self._versions_seen:
            # kernel-by-kernel iteration: first sight of a configuration, or parameters changed from outside
            # (load_state_dict ...) so the bf16 weight images must be refreshed by the host-side cache logic
            buf = self._buffers(B, R, dev) if ent is None else ent.buf
            self._load(buf, images, z1, z2, eps, z3)
            stats = self._run_eager(buf)
            self._versions_seen = self._versions()
            if self.use_graph is not False and ent is None and key == self._last_key:
                self._capture(key, B, R, dev)
            self._last_key = key
            self.last_run = 'eager'
            return stats.clone() if ent is not None else stats
        self._graphs[key] = self._graphs.pop(key)                 # most recently used last
        self._load(ent.buf, images, z1, z2, eps, z3)
        self.opt_d.advance()
        self.opt_g.advance()
        self.inputs_loaded = torch.cuda.Event()
        self.inputs_loaded.record()
        if len(ent.graphs) == 1:
            ent.graphs[0].replay()
        else:
            ev = None
            if self.segment_events is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                self.segment_events.append(ev)
                ev[0].record()
            cur = torch.cuda.current_stream()
            ent.graphs[0].replay()
            if self.dp:
                # the critic's 2 MB all-reduce runs on the communication stream beside the generator's forward pass
                if self._comm_stream is None:
                    self._comm_stream = torch.cuda.Stream()
                self._comm_stream.wait_stream(cur)
                with torch.cuda.stream(self._comm_stream):
                    self._allreduce(ent.flats[0])
            rest = ent.graphs[1:]
            if len(rest) == 3:
                rest.pop(0).replay()    # the generator's forward pass, beside the all-reduce
            if self.dp:
                cur.wait_stream(self._comm_stream)
            if ev:
                ev[1].record()
            rest[0].replay()
            self._exchange_g(ent.buf)
            if ev:
                ev[2].record()
            rest[1].replay()
            if ev:
                ev[3].record()
        self._last_key = key
        self.last_run = 'replay'
        self.launches_per_step = ent.launches
        _lib.launch_count += ent.launches
        return ent.stats.clone()

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
