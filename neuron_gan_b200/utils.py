"""The subset of the reference's utils.py that the hot path and its callers import by name
(train.py:20-21, loss_functions.py:3, eval.py:6): latent sampling, checkpointing, sample generation.
Plotting / video / memory logging (reference utils.py:228-343, 360-543, 614-788) are out of scope."""
import datetime
import os
import pickle

import numpy as np
import torch
from torch import nn

# ----------------------------------------------------------------------------------------------- training
Latent_vecs_memo = {}


def sample_latent_vec(size: tuple, seed=None, mode='randn', device=torch.device('cpu')):
    """Latent batch, drawn on the CPU generator for reproducibility and then moved (reference utils.py:57-92):
    'randn' -> clamp(+-5) -> rows L2-normalised (uniform on the hypersphere); 'rand' -> U(-1, 1).
    With `seed`, the global RNG state is saved/restored and the result memoised by (size, mode, seed)."""
    device = torch.device(device)
    if seed is not None:
        key = (size, mode, seed)
        if key in Latent_vecs_memo:
            return Latent_vecs_memo[key].to(device)
        rng_state = torch.get_rng_state()
        torch.manual_seed(seed)
    if mode == 'rand':
        z = 2 * torch.rand(*size, device='cpu') - 1
    elif mode == 'randn':
        z = torch.randn(*size, device='cpu').clamp(-5, 5)
        z = z / z.norm(p=2, dim=1, keepdim=True)
    else:
        raise ValueError('{} is not supported'.format(mode))
    if seed is not None:
        torch.set_rng_state(rng_state)
        Latent_vecs_memo[key] = z
    if device.type != 'cpu':
        z = z.to(device)
    return z


def get_saved_attrs(model):
    return {a: getattr(model, a) for a in getattr(model, 'saved_attrs', [])}


def set_saved_attrs(model, saved_attrs_dict):
    for name, value in saved_attrs_dict.items():
        if hasattr(model, name):
            setattr(model, name, value)
        else:
            raise ValueError('{} is not an attribute of {}', name, model)
    return saved_attrs_dict


class Checkpointer:
    """Reference-format .pth checkpoints (reference utils.py:124-223): keys epoch, Generator_state,
    Generator_attrs, Discriminator_state, Discriminator_attrs, lr, Loss_real, Loss_fake, Loss_G, Loss_D.
    No optimiser state, like the reference.  `load_state(filename)` reads the weights from `filename`
    (the reference re-reads self.filename there, utils.py:213-215, which is a bug)."""

    def __init__(self, Generator_net, Discriminator_net, lr: float, filename: str, N_epochs=100, verbose=True,
                 device=torch.device('cpu'), extra_checkpoint_period=50e3):
        self.Generator_net = Generator_net
        self.Discriminator_net = Discriminator_net
        self.lr = lr
        self.filename = filename
        self.epoch = 0
        self.Loss_real = np.zeros(N_epochs)
        self.Loss_fake = np.zeros(N_epochs)
        self.Loss_G = np.zeros(N_epochs)
        self.Loss_D = np.zeros(N_epochs)
        self.verbose = verbose
        self.device = device
        self.extra_checkpoint_period = extra_checkpoint_period

    def save_state(self, epoch):
        self.epoch = epoch
        ckpt = {'epoch': self.epoch,
                'Generator_state': self.Generator_net.state_dict(),
                'Generator_attrs': get_saved_attrs(self.Generator_net),
                'Discriminator_state': self.Discriminator_net.state_dict(),
                'Discriminator_attrs': get_saved_attrs(self.Discriminator_net),
                'lr': self.lr,
                'Loss_real': self.Loss_real[:epoch], 'Loss_fake': self.Loss_fake[:epoch],
                'Loss_G': self.Loss_G[:epoch], 'Loss_D': self.Loss_D[:epoch]}
        torch.save(ckpt, self.filename)
        if epoch % self.extra_checkpoint_period == 0:
            base, ext = os.path.splitext(self.filename)
            torch.save(ckpt, base + '_{:d}k'.format(int(epoch / 1000)) + ext)
        if self.verbose:
            print('Training state at epoch {} saved in {}.'.format(self.epoch, self.filename))

    def load_state(self, filename=None):
        source = self.filename if filename is None else filename
        ckpt = torch.load(source, map_location=self.device, weights_only=False)
        if filename is None:
            self.epoch = ckpt['epoch']
            self.Loss_real[:self.epoch] = ckpt['Loss_real']
            self.Loss_fake[:self.epoch] = ckpt['Loss_fake']
            self.Loss_G[:self.epoch] = ckpt['Loss_G']
            self.Loss_D[:self.epoch] = ckpt['Loss_D']
        if 'Generator_attrs' in ckpt and 'Discriminator_attrs' in ckpt:
            gen_attrs = {k: v for k, v in ckpt['Generator_attrs'].items() if k in self.Generator_net.saved_attrs}
            dis_attrs = {k: v for k, v in ckpt['Discriminator_attrs'].items()
                         if k in self.Discriminator_net.saved_attrs}
            if hasattr(self.Generator_net, 'set_resolution'):
                res, alpha = gen_attrs['image_size'], float(gen_attrs['alpha'])
                self.Generator_net.set_resolution(res, alpha)
                self.Discriminator_net.set_resolution(res, alpha)
            dev_g = next(self.Generator_net.parameters()).device
            for attrs in (gen_attrs, dis_attrs):
                if torch.is_tensor(attrs.get('alpha')):
                    attrs['alpha'] = attrs['alpha'].to(dev_g)
            set_saved_attrs(self.Generator_net, gen_attrs)
            set_saved_attrs(self.Discriminator_net, dis_attrs)
        gen = self.Generator_net.from_state_dict(source, device=torch.device('cpu'), verbose=False)
        dis = self.Discriminator_net.from_state_dict(source, device=torch.device('cpu'), verbose=False)
        self.Generator_net.load_state_dict(gen.state_dict(), strict=False)
        self.Discriminator_net.load_state_dict(dis.state_dict(), strict=False)
        if self.verbose and filename is None:
            print('Loaded training state from {}'.format(self.filename))
        elif self.verbose:
            print('Loaded weights from {}'.format(filename))


def init_weights(m: nn.Module):
    """DCGAN initialisation (reference utils.py:96-101); train.py:208-209 applies it to the legacy nets only."""
    if type(m) in [nn.Conv2d, nn.ConvTranspose2d]:
        m.weight.data.normal_(0.0, 0.02)
    elif type(m) == nn.BatchNorm2d:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0.0)


def ValidatedInput(prompt: str, validate_func, invalid_ans_msg='Invalid answer.'):
    """Interactive prompt of the reference (utils.py:234-246; train.py:117-125, configs/config.py:138-144)."""
    if not prompt.endswith('\n'):
        prompt += '\n'
    while True:
        Ans = input(prompt)
        if validate_func(Ans):
            break
        print(invalid_ans_msg)
    return Ans


def N_params(model: nn.Module):
    return sum(p.numel() for p in model.parameters())


def calculate_grad_norm_hist(model: nn.Module, grad_min=-30, log_scale=True):
    """|grad| of every parameter element that has one, log10-scaled (reference utils.py:249-275)."""
    parts = [p.grad.detach().abs().flatten() for p in model.parameters() if p.grad is not None and p.requires_grad]
    vals = torch.cat(parts).cpu().numpy() if parts else np.array([])
    if log_scale:
        vals = np.log10(np.maximum(vals, 10.0 ** grad_min))
    else:
        vals = np.maximum(vals, grad_min)
    if vals.size:
        return vals, float(np.mean(vals)), float(np.std(vals))
    return vals, float('nan'), float('nan')


def _pyplot():
    try:
        import matplotlib
        matplotlib.use('Agg')
        import matplotlib.pyplot as plt
        return plt if hasattr(plt, 'subplots') and callable(getattr(plt, 'subplots')) else None
    except Exception:  # noqa: BLE001 - matplotlib is optional: the plots are observability only
        return None


def plot_grad_norm(generator_model, discriminator_model, filename: str = None):
    """Histogram of the parameter-gradient magnitudes of both networks (reference utils.py:619-645).  Needs real fp32
    `.grad` tensors on the parameters, which both the autograd path and TrainStep provide.  Without matplotlib the
    statistics are still computed (and returned); no figure is written."""
    g_vals, g_mean, g_std = calculate_grad_norm_hist(generator_model)
    d_vals, d_mean, d_std = calculate_grad_norm_hist(discriminator_model)
    plt = _pyplot()
    if plt is not None and filename is not None:
        try:
            fig, (ax1, ax2) = plt.subplots(1, 2, figsize=(8, 5))
            ax1.hist(g_vals, alpha=0.75)
            ax1.set_title('Generator, mean={:.2}, std={:.2}'.format(g_mean, g_std))
            ax2.hist(d_vals, alpha=0.75)
            ax2.set_title('Discriminator, mean={:.2}, std={:.2}'.format(d_mean, d_std))
            for ax in (ax1, ax2):
                ax.set_xlabel('Parameter gradient norm (Logged)')
                ax.set_ylabel('Counts')
            fig.tight_layout()
            fig.savefig(filename)
            plt.close(fig)
        except Exception:  # noqa: BLE001 - a stubbed matplotlib draws nothing
            pass
    return (g_mean, g_std), (d_mean, d_std)


def plot_scores(loss_real, loss_fake, filename, G_loss=None, D_loss=None):
    """Training summary plot (reference utils.py:649-665); a no-op without matplotlib."""
    plt = _pyplot()
    if plt is None:
        return
    try:
        fig = plt.figure()
        plt.plot(loss_real, label='Real images (<D(x)>_x)')
        plt.plot(loss_fake, label='Fake images (<D(G(z))>_z)')
        if G_loss:
            plt.plot(G_loss, label='Generator')
        if D_loss:
            plt.plot(D_loss, label='Discriminator')
        plt.legend(loc='upper left')
        plt.xlabel('Epoch')
        plt.savefig(filename)
        plt.close(fig)
    except Exception:  # noqa: BLE001
        pass


def save_vars(variables: dict, directory='./saved_vars', verbose=True):
    """NaN-guard dump (reference utils.py:308-342): pickles the tensors / numbers found in `variables`."""
    os.makedirs(directory, exist_ok=True)
    stamp = datetime.datetime.now().strftime('%Y%m%d_%H%M%S')
    path = os.path.join(directory, f'vars_{stamp}.pkl')
    keep = {}
    for k, v in variables.items():
        if torch.is_tensor(v):
            keep[k] = v.detach().cpu()
        elif isinstance(v, (int, float, str, np.ndarray)):
            keep[k] = v
    with open(path, 'wb') as f:
        pickle.dump(keep, f)
    if verbose:
        print(f'Variables saved in:\n{path}')
    return path


def Calculate_D_steps(Loss_real, Loss_fake, N_min, N_max, Period):
    """`adapt_critic` (reference utils.py:105-120, called at train.py:336-338): critic steps for the next epoch,
    round(N_max * std(real scores) / mean|fake - real|) over the last `Period` iterations, clipped to [N_min, N_max];
    N_max while the series are still empty.  Assign the result to `TrainStep.n_critic` between iterations."""
    if not (Loss_real and Loss_fake):
        return N_max
    real = np.asarray(Loss_real[-Period:])
    fake = np.asarray(Loss_fake[-Period:])
    spread = np.std(real)
    gap = np.mean(np.abs(fake - real))
    return int(max(min(np.round(spread / gap * N_max), N_max), N_min))


# ----------------------------------------------------------------------------------------------- testing
def gen_samples(Generator: nn.Module, N_images=16, seed=None, chunk=128, dtype=torch.float32):
    """Generator-only inference (reference utils.py:346-355): returns (images [N, 1, R, R] on the generator's device,
    z).  Chunked so that 4096 samples at 512x512 fit; every chunk writes straight into its slice of the result (no
    concatenation, no per-chunk allocation of the result).  dtype: torch.float32 (reference) or torch.bfloat16."""
    from . import engine
    dev = next(Generator.parameters()).device
    if dev.type != 'cuda':
        raise RuntimeError('neuron_gan_b200 generators run on CUDA only: move the network to the GPU first')
    z_latent = sample_latent_vec((N_images, Generator.latent_dim), seed=seed, device=dev)
    R = Generator.image_size
    images = torch.empty((N_images, 1, R, R), dtype=dtype, device=dev)
    with torch.no_grad():
        for i in range(0, N_images, chunk):
            engine.g_forward(Generator, z_latent[i:i + chunk], save=False, img_out=images[i:i + chunk, 0])
    return images, z_latent


_host_pool = {}


def gen_samples_host(Generator: nn.Module, N_images=16, seed=None, dtype=torch.float32, chunk=128, to_host=True):
    """eval.py's data path end to end: seeded latents -> generator -> images in (pinned) HOST memory, which is where
    the reference has them before it writes the PNG grid (utils.py:583 `.cpu()`).  Chunks are double-buffered on
    the device and their device-to-host copies run on a copy stream beside the next chunk's kernels; the pinned
    result buffer is kept per (N, R, dtype).  Returns the host tensor [N, 1, R, R] (valid until the next call with
    the same shape), or -- with to_host=False -- the last device chunk (timing without the copy)."""
    from . import engine
    dev = next(Generator.parameters()).device
    R = Generator.image_size
    z_latent = sample_latent_vec((N_images, Generator.latent_dim), seed=seed, device=dev)
    key = (N_images, R, dtype, dev, min(chunk, N_images))
    pool = _host_pool.get(key)
    if pool is None:
        if len(_host_pool) >= 4:
            _host_pool.clear()
        c = min(chunk, N_images)
        pool = _host_pool[key] = {
            'host': torch.empty((N_images, 1, R, R), dtype=dtype).pin_memory(),
            'dev': [torch.empty((c, R, R), dtype=dtype, device=dev) for _ in range(2)],
            'free': [None, None], 'stream': torch.cuda.Stream(dev)}
    host, bufs, free, copy_stream = pool['host'], pool['dev'], pool['free'], pool['stream']
    cur = torch.cuda.current_stream(dev)
    with torch.no_grad():
        for k, i in enumerate(range(0, N_images, chunk)):
            n = min(chunk, N_images - i)
            buf = bufs[k % 2][:n]
            if free[k % 2] is not None:
                cur.wait_event(free[k % 2])              # the copy that last read this buffer has finished
            engine.g_forward(Generator, z_latent[i:i + n], save=False, img_out=buf)
            if to_host:
                ready = torch.cuda.Event()
                ready.record(cur)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ready)
                    host[i:i + n, 0].copy_(buf, non_blocking=True)
                    free[k % 2] = torch.cuda.Event()
                    free[k % 2].record(copy_stream)
    if not to_host:
        return buf
    cur.wait_stream(copy_stream)                         # stream-ordered: callers synchronise before reading `host`
    return host


def plot_gen_samples(Generator: nn.Module, eval_noise=None, N_images=16, seed=None, filename=None):
    """Sample grid PNG (reference utils.py:568-610): nearest-upsample to image_size_max, nrow=round(sqrt(N)),
    normalize=True.  Needs `filename` (the reference's matplotlib display path is out of scope)."""
    if isinstance(eval_noise, int):
        # eval.py:26 passes `n` positionally, i.e. into this slot (as shipped the reference
        # then calls Generator(20) and raises); the intended call is plot_gen_samples(G, N_images=n, filename=...)
        N_images, eval_noise = eval_noise, None
    was_training = Generator.training
    Generator.train(False)
    if eval_noise is None:
        images = gen_samples_host(Generator, N_images, seed=seed)
        torch.cuda.synchronize()
        images = images.clone()
    else:
        with torch.no_grad():
            images = Generator(eval_noise).detach().cpu()
        N_images = images.size(0)
    Generator.train(was_training)
    n_rows = int(np.round(np.sqrt(N_images)))
    if images.size(-1) != Generator.image_size_max:
        size = (Generator.image_size_max, Generator.image_size_max)
        images = nn.functional.interpolate(images, size=size)
    if filename is None:
        raise ValueError('plot_gen_samples needs a filename (interactive display is not part of this build)')
    import torchvision
    torchvision.utils.save_image(images, filename, nrow=n_rows, normalize=True)
    return images


class DevicePrefetcher:
    """Iterate host batches (ideally pinned) as device tensors with the next batch's host-to-device copy running on
    its own stream while the current one is consumed -- the role `pin_memory=True` + `non_blocking` copies play in
    the reference's loop (train.py:150-154, 352-353).  The yielded tensor is only valid until the next iteration.
    `batches` is an iterable, or a callable returning one (so the object can be iterated once per epoch).

    `gate`: optional callable returning a CUDA event (or None) that the NEXT batch's copy must wait for.  The copy
    engine is shared: a 17 MB image copy that becomes runnable at the same moment as the consumer's own small
    per-iteration copies (latent draws, optimiser scalars) delays them -- and the iteration -- by its full 0.7 ms.
    Passing `gate=lambda: step.inputs_loaded` orders the big copy behind them; it then runs beside the kernels."""

    def __init__(self, batches, device, depth=2, gate=None):
        self.batches, self.device, self.depth = batches, torch.device(device), max(2, int(depth))
        self.gate = gate
        self._stream = None
        self._slots = [None] * self.depth          # device staging, kept across iterations of this object

    def __iter__(self):
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        copy_stream = self._stream
        slots, ready, free = self._slots, [None] * self.depth, [None] * self.depth
        it = iter(self.batches() if callable(self.batches) else self.batches)

        def launch(k, host, after=None):
            with torch.cuda.stream(copy_stream):
                if free[k] is not None:
                    copy_stream.wait_event(free[k])          # the consumer has finished reading this slot
                if after is not None:
                    copy_stream.wait_event(after)
                if slots[k] is None or slots[k].shape != host.shape or slots[k].dtype != host.dtype:
                    slots[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
                slots[k].copy_(host, non_blocking=True)
                ready[k] = torch.cuda.Event()
                ready[k].record(copy_stream)

        first = next(it, None)
        if first is None:
            return
        launch(0, first)
        i = 0
        while True:
            k = i % self.depth
            torch.cuda.current_stream(self.device).wait_event(ready[k])
            yield slots[k]
            # the consumer has queued its work on this batch: release the slot behind it and only now queue the next
            # batch's copy (behind the consumer's own copies of this iteration, see `gate`)
            free[k] = torch.cuda.Event()
            free[k].record(torch.cuda.current_stream(self.device))
            nxt = next(it, None)
            if nxt is None:
                return
            launch((i + 1) % self.depth, nxt, self.gate() if self.gate is not None else None)
            i += 1


class PinnedRing:
    """A few sets of pinned host staging buffers reused round-robin.  Asynchronous host-to-device copies need pinned
    sources that stay untouched until the copy has run; allocating fresh pinned memory every step costs a
    cudaHostAlloc whenever the caching allocator has nothing free (milliseconds of jitter), so the buffers are
    allocated once and each set is guarded by an event recorded after its last use."""

    def __init__(self, shapes, dtype=torch.float32, depth=4):
        self.sets = [[torch.empty(s, dtype=dtype).pin_memory() for s in shapes] for _ in range(depth)]
        self.events = [None] * depth
        self.i = 0

    def acquire(self):
        k = self.i % len(self.sets)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()          # long complete unless the host runs > depth steps ahead
        return k, self.sets[k]

    def release(self, k):
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
