"""CPU oracle for neuron-gan's progressive-growing WGAN-GP training step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``neuron_gan_b200``) may import,
call or execute this module; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of
``bench.py`` (``cpu_baseline`` / ``--impl reference``) do, and there only as the checker / the
CPU baseline.

What it is: a functional restatement (plain PyTorch on CPU, fp32 or fp64) of the reference's
algorithm for the hot path, each function citing the reference ``file:line`` it follows.  The
arithmetic itself lives in a third-party dependency of the reference -- PyTorch
(``requirements.txt``: ``torch==1.13.1``; this image has torch 2.11) -- so the restatement calls the
same ATen operators (``conv2d``, ``interpolate``, ``avg_pool2d``, ``autograd.grad``) in the same order.

Pinning: the reference ships no tests and no golden vectors ("parity unpinned" by the reference's
own tests).  The oracle is therefore pinned against outputs of the reference
itself, run in the build container: ``tests/golden/gen_golden.py`` imports ``/root/reference`` and
writes ``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` checks this file against them on every
CPU test run, and ``tests/test_oracle_vs_reference.py`` compares live when ``/root/reference`` exists.

Parameter naming used here ("level" = index into the feature list, level 0 of G is the lowest
resolution, level L-1 of D is the lowest resolution):

  G: ``lin.w`` [F0*S*S, latent], ``conv0.w`` [F0,F0,3,3], ``blk{i}.w1`` [F_i,F_{i-1},3,3],
     ``blk{i}.w2`` [F_i,F_i,3,3] for i=1..L-1, ``toim{i}.w`` [1,F_i,1,1] for i=0..L-1
  D: ``from{i}.w`` [F_i,1,1,1], ``from{i}.b`` [F_i] for i=0..L-1, ``blk{i}.w1`` [F_{i+1},F_i,3,3],
     ``blk{i}.w2`` [F_{i+1},F_{i+1},3,3] for i=0..L-2, ``last.w`` [F,F,3,3], ``last.b``,
     ``head.w`` [1,F,S,S], ``head.b``

``g_state_to_named`` / ``d_state_to_named`` translate the reference's ``state_dict()`` keys (which are
renumbered at every resolution transition, models.py:374-377, 546-549) to these names.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# configuration (configs/config.py:58-63, train.py:162-165)
# ----------------------------------------------------------------------------------------------


@dataclass
class Arch:
    gen_features: list = field(default_factory=lambda: [128, 64, 32, 32, 16, 16])
    dis_features: list = field(default_factory=lambda: [16, 16, 32, 32, 64, 128])
    latent_dim: int = 512
    image_size: int = 512
    n_colors: int = 1
    leak: float = 0.2

    @property
    def size_init(self) -> int:  # train.py:162-165
        return self.image_size // (2 ** (len(self.gen_features) - 1))


def he_gain(leak: float) -> float:
    """torch.nn.init.calculate_gain('leaky_relu', leak) (models.py:198, 235)."""
    return math.sqrt(2.0 / (1.0 + leak * leak))


def conv_scale(weight: torch.Tensor, leak: float) -> float:
    """Runtime equalised-LR scale, fan_in mode (models.py:186-201, 224-238).

    The reference stores it as a 0-dim buffer; multiplied into an fp32 activation it acts as the
    fp32-rounded scalar."""
    fan_in = weight[0].numel()
    return he_gain(leak) / math.sqrt(fan_in)


# ----------------------------------------------------------------------------------------------
# optional emulation of the CUDA path's storage precision (test aid, off by default)
# ----------------------------------------------------------------------------------------------
_EMULATE_BF16 = False


class emulate_bf16:
    """Context manager: round conv/linear weights and every stored feature map to bf16 (straight-through in
    backward), i.e. the places where the CUDA path keeps bf16 in HBM.  With it the oracle reproduces the
    LeakyReLU-mask flips that bf16 operands cause, so end-to-end GRADIENTS of the CUDA path can be compared
    at a few-percent tolerance; without it only losses and per-layer outputs are comparable.
    The reference itself is fp32: this mode is never used for the reference-pinned golden checks."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        global _EMULATE_BF16
        self.prev, _EMULATE_BF16 = _EMULATE_BF16, self.on

    def __exit__(self, *exc):
        global _EMULATE_BF16
        _EMULATE_BF16 = self.prev


def _q(x):
    if not _EMULATE_BF16:
        return x
    return x + (x.detach().to(torch.bfloat16).to(x.dtype) - x.detach())


# ----------------------------------------------------------------------------------------------
# optional frozen LeakyReLU masks (test aid, off by default)
# ----------------------------------------------------------------------------------------------
_FORCED_MASKS = None


class forced_masks:
    """Context manager: every LeakyReLU inside takes its mask (True = positive side) from `masks`, in call order,
    instead of from the sign of its input.  LeakyReLU is the only non-smooth operation of the networks; at random
    initialisation the outputs are random-sign sums over pixels, so the ~0.3 % of masks that two bf16 evaluations
    decide differently move a gradient by ~sqrt(0.3 %) -- that conditioning, not kernel error, is what limits a
    free-running comparison at 256x256 / 512x512.  With the CUDA path's own masks (recovered from the signs of its
    saved activations) the oracle is a smooth function of everything else, and whole-network gradients agree to the
    accumulated rounding error: a composition error in any layer shows at full size."""

    def __init__(self, masks):
        self.masks = list(masks)

    def __enter__(self):
        global _FORCED_MASKS
        self.prev, _FORCED_MASKS = _FORCED_MASKS, iter(self.masks)
        return self

    def __exit__(self, *exc):
        global _FORCED_MASKS
        left = sum(1 for _ in _FORCED_MASKS)
        _FORCED_MASKS = self.prev
        if exc[0] is None and left:
            raise AssertionError(f'forced_masks: {left} masks were not consumed')


def lrelu(x, leak):
    """nn.LeakyReLU (models.py:263, 267, 310, 315, 472)."""
    if _FORCED_MASKS is None:
        return F.leaky_relu(x, leak)
    m = next(_FORCED_MASKS)
    assert m.shape == x.shape, (m.shape, x.shape)
    return x * torch.where(m, torch.ones((), dtype=x.dtype), torch.full((), leak, dtype=x.dtype))


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------


def pixel_norm(x: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """models.py:110-126 (option 2): x / sqrt(mean_c(x^2) + eps)."""
    return x / torch.sqrt(torch.mean(x ** 2, dim=1, keepdim=True) + eps)


def eq_conv(x, w, b=None, pad=1, leak=0.2):
    """Conv2d_normalized.forward (models.py:203-204): conv(scale * x, W) + b; bias is not scaled."""
    s = torch.tensor(conv_scale(w, leak), dtype=x.dtype)
    if _EMULATE_BF16 and pad == 1:         # the 3x3 convs run on bf16 operands; the SxS head keeps fp32 weights
        w = _q(w)
    return F.conv2d(s * x, w, b, stride=1, padding=pad)


def eq_linear(z, w, leak=0.2):
    """Linear_normalized.forward (models.py:240-241), no bias in G (models.py:299-300)."""
    s = torch.tensor(he_gain(leak) / math.sqrt(w.shape[1]), dtype=z.dtype)
    if _EMULATE_BF16:                       # the CUDA GEMM takes both operands in bf16 and scales the fp32 sums
        return s * F.linear(_q(z), _q(w))
    return F.linear(s * z, w)


def up2(x):
    """Interpolate(scale_factor=2, mode='bilinear') (models.py:78-89, 257, 335)."""
    y = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=None)
    return _q(y) if x.shape[1] > 1 else y      # feature maps are stored in bf16, 1-channel images in fp32


def down2_bilinear(x):
    """Interpolate(scale_factor=0.5, mode='bilinear') (models.py:507); equals 2x2 mean."""
    return F.interpolate(x, scale_factor=0.5, mode='bilinear', align_corners=None)


def clp(x, w, b, leak):
    """conv -> LeakyReLU -> PixelNorm (models.py:261-268, 312-316, 469-473)."""
    return _q(pixel_norm(lrelu(eq_conv(x, w, b, 1, leak), leak)))


def scale_block(x, w1, w2, up: bool, leak):
    """Conv2d_scale_block (models.py:245-268): resample first, then two conv/lrelu/PN stages."""
    x = up2(x) if up else _q(F.avg_pool2d(x, 2))
    return clp(clp(x, w1, None, leak), w2, None, leak)


def to_image(x, w):
    """ToImage (models.py:133-149): plain 1x1 conv, no scale, no bias, then tanh."""
    return torch.tanh(F.conv2d(x, w))


def from_image(x, w, b):
    """FromImage (models.py:156-165): plain 1x1 conv with bias, no activation."""
    return F.conv2d(x, w, b)


# ----------------------------------------------------------------------------------------------
# networks
# ----------------------------------------------------------------------------------------------


def g_forward(p: dict, z, n_layers: int, alpha: float, arch: Arch):
    """Generator_PG.forward (models.py:344-353) at ``n_layers`` resolution levels.

    alpha < 1: levels 0..n-2 are the stable trunk, level n-1 is being faded in."""
    leak, s0, f0 = arch.leak, arch.size_init, arch.gen_features[0]
    x = eq_linear(z, p['lin.w'], leak).unflatten(1, (f0, s0, s0))          # models.py:299-302
    x = _q(pixel_norm(lrelu(x, leak)))                                      # models.py:310-311
    x = clp(x, p['conv0.w'], None, leak)                                    # models.py:312-316
    n_trunk = n_layers - 1 if alpha >= 1 else n_layers - 2
    for i in range(1, n_trunk + 1):
        x = scale_block(x, p[f'blk{i}.w1'], p[f'blk{i}.w2'], True, leak)
    if alpha >= 1:
        return to_image(x, p[f'toim{n_trunk}.w'])                           # models.py:353
    i = n_layers - 1
    im_start = up2(to_image(x, p[f'toim{i - 1}.w']))                        # models.py:348
    im_end = to_image(scale_block(x, p[f'blk{i}.w1'], p[f'blk{i}.w2'], True, leak), p[f'toim{i}.w'])
    return im_start + alpha * (im_end - im_start)                           # models.py:350


def d_forward(p: dict, x, n_layers: int, alpha: float, arch: Arch):
    """Discriminator_PG.forward (models.py:516-524). Level L-1 is the lowest resolution."""
    leak, L = arch.leak, len(arch.dis_features)
    top = L - n_layers                      # level whose FromImage sees the full-resolution input
    if alpha >= 1:
        y = from_image(x, p[f'from{top}.w'], p[f'from{top}.b'])             # models.py:524
        first_blk = top
        if first_blk == L - 1:
            y = _q(y)         # (bf16 emulation only) no block follows: FromImage's output is what gets stored
    else:
        y_start = from_image(down2_bilinear(x), p[f'from{top + 1}.w'], p[f'from{top + 1}.b'])
        y_end = scale_block(from_image(x, p[f'from{top}.w'], p[f'from{top}.b']),
                            p[f'blk{top}.w1'], p[f'blk{top}.w2'], False, leak)
        y = _q(y_start + alpha * (y_end - y_start))                         # models.py:519-521
        first_blk = top + 1
    for i in range(first_blk, L - 1):
        y = scale_block(y, p[f'blk{i}.w1'], p[f'blk{i}.w2'], False, leak)
    y = clp(y, p['last.w'], p['last.b'], leak)                              # models.py:469-473
    y = eq_conv(y, p['head.w'], p['head.b'], 0, leak)                       # models.py:485-487
    return y.flatten(1)                                                     # models.py:490


# ----------------------------------------------------------------------------------------------
# parameter construction in the reference's RNG order
# ----------------------------------------------------------------------------------------------


def _kaiming(m, leak):
    """kaiming_init (models.py:31-34)."""
    nn.init.kaiming_normal_(m.weight, a=leak, mode='fan_in', nonlinearity='leaky_relu')
    if m.bias is not None:
        m.bias.data.zero_()
    return m


def build_g_params(arch: Arch) -> "OrderedDict[str, torch.Tensor]":
    """Draw G's parameters exactly as Generator_PG.__init__ does (models.py:295-329).

    nn.Linear / nn.Conv2d constructors consume RNG for their default init before kaiming_init
    overwrites it, so the torch constructors are called here in the same order."""
    f, s0, leak = arch.gen_features, arch.size_init, arch.leak
    p = OrderedDict()
    p['lin.w'] = _kaiming(nn.Linear(arch.latent_dim, f[0] * s0 * s0, bias=False), leak).weight.data
    p['conv0.w'] = _kaiming(nn.Conv2d(f[0], f[0], 3, padding=1, bias=False), leak).weight.data
    for i in range(1, len(f)):
        p[f'blk{i}.w1'] = _kaiming(nn.Conv2d(f[i - 1], f[i], 3, padding=1, bias=False), leak).weight.data
        p[f'blk{i}.w2'] = _kaiming(nn.Conv2d(f[i], f[i], 3, padding=1, bias=False), leak).weight.data
    for i in range(len(f)):
        p[f'toim{i}.w'] = _kaiming(nn.Conv2d(f[i], arch.n_colors, 1, bias=False), leak).weight.data
    return p


def build_d_params(arch: Arch) -> "OrderedDict[str, torch.Tensor]":
    """Draw D's parameters exactly as Discriminator_PG.__init__ does (models.py:468-503)."""
    f, s0, leak = arch.dis_features, arch.size_init, arch.leak
    p = OrderedDict()
    m = _kaiming(nn.Conv2d(f[-1], f[-1], 3, padding=1), leak)
    p['last.w'], p['last.b'] = m.weight.data, m.bias.data
    m = _kaiming(nn.Conv2d(f[-1], 1, (s0, s0), padding=0), leak)
    p['head.w'], p['head.b'] = m.weight.data, m.bias.data
    for i in range(len(f) - 1):
        p[f'blk{i}.w1'] = _kaiming(nn.Conv2d(f[i], f[i + 1], 3, padding=1, bias=False), leak).weight.data
        p[f'blk{i}.w2'] = _kaiming(nn.Conv2d(f[i + 1], f[i + 1], 3, padding=1, bias=False), leak).weight.data
    for i in range(len(f)):
        m = _kaiming(nn.Conv2d(arch.n_colors, f[i], 1), leak)
        p[f'from{i}.w'], p[f'from{i}.b'] = m.weight.data, m.bias.data
    return p


def build_params(arch: Arch, seed: int = 1):
    """torch.manual_seed(seed) -> G -> D, the order of train.py:114, 172, 184."""
    torch.manual_seed(seed)
    g = build_g_params(arch)
    d = build_d_params(arch)
    return g, d


# ----------------------------------------------------------------------------------------------
# state-dict key translation (models.py:355-392, 526-564)
# ----------------------------------------------------------------------------------------------


def n_layers_for(res: int, arch: Arch) -> int:
    return int(round(math.log2(res / arch.size_init))) + 1


def g_key_map(n_layers: int, in_transition: bool, arch: Arch) -> dict:
    """oracle name -> reference state_dict key for a generator at this structural state."""
    L = len(arch.gen_features)
    n_trunk = n_layers - 2 if in_transition else n_layers - 1
    km = {'lin.w': 'layers.0.weight', 'conv0.w': 'layers.4.weight'}
    for i in range(1, L):
        if i <= n_trunk:
            base = f'layers.{6 + i}'
        else:
            base = f'conv_block_list.{i - n_trunk - 1}'
        km[f'blk{i}.w1'], km[f'blk{i}.w2'] = base + '.1.weight', base + '.4.weight'
    for i in range(L):
        if i < n_trunk:
            continue                      # popped and dropped: no longer in the state dict
        km[f'toim{i}.w'] = 'ToIm.layers.0.weight' if i == n_trunk else f'ToIm_list.{i - n_trunk - 1}.layers.0.weight'
    return km


def d_key_map(n_layers: int, in_transition: bool, arch: Arch) -> dict:
    """oracle name -> reference state_dict key for a discriminator at this structural state."""
    L = len(arch.dis_features)
    n_trunk = n_layers - 2 if in_transition else n_layers - 1      # blocks moved into `layers`
    first_trunk_blk = L - 1 - n_trunk
    km = {'last.w': f'layers.{n_trunk}.weight', 'last.b': f'layers.{n_trunk}.bias',
          'head.w': f'layers.{n_trunk + 3}.weight', 'head.b': f'layers.{n_trunk + 3}.bias'}
    for i in range(L - 1):
        base = f'layers.{i - first_trunk_blk}' if i >= first_trunk_blk else f'conv_block_list.{i}'
        km[f'blk{i}.w1'], km[f'blk{i}.w2'] = base + '.1.weight', base + '.4.weight'
    cur = L - 1 - n_trunk                                           # level of the current FromIm
    for i in range(L):
        if i > cur:
            continue
        base = 'FromIm.conv' if i == cur else f'FromIm_list.{i}.conv'
        km[f'from{i}.w'], km[f'from{i}.b'] = base + '.weight', base + '.bias'
    return km


def g_state_to_named(state: dict, n_layers: int, in_transition: bool, arch: Arch) -> dict:
    return {k: state[v] for k, v in g_key_map(n_layers, in_transition, arch).items()}


def d_state_to_named(state: dict, n_layers: int, in_transition: bool, arch: Arch) -> dict:
    return {k: state[v] for k, v in d_key_map(n_layers, in_transition, arch).items()}


def active_g_names(n_layers: int, alpha: float, arch: Arch) -> list:
    """Parameters that receive a gradient at this state (everything else has grad None)."""
    names = ['lin.w', 'conv0.w']
    for i in range(1, n_layers):
        names += [f'blk{i}.w1', f'blk{i}.w2']
    names.append(f'toim{n_layers - 1}.w')
    if alpha < 1:
        names.append(f'toim{n_layers - 2}.w')
    return names


def active_d_names(n_layers: int, alpha: float, arch: Arch) -> list:
    L = len(arch.dis_features)
    top = L - n_layers
    names = ['last.w', 'last.b', 'head.w', 'head.b', f'from{top}.w', f'from{top}.b']
    for i in range(top, L - 1):
        names += [f'blk{i}.w1', f'blk{i}.w2']
    if alpha < 1:
        names += [f'from{top + 1}.w', f'from{top + 1}.b']
    return names


# ----------------------------------------------------------------------------------------------
# latent sampling and losses
# ----------------------------------------------------------------------------------------------


def sample_latent(size, generator=None):
    """sample_latent_vec, mode='randn' (utils.py:57-92): CPU randn, clamp +-5, L2-normalise rows."""
    z = torch.randn(*size, device='cpu', generator=generator).clamp(-5, 5)
    return z / z.norm(p=2, dim=1, keepdim=True)


def d_w_loss(gp, dp, x, z, n_layers, alpha, arch, drift=0.0):
    """D_W_loss.forward (loss_functions.py:14-47). Returns (D_loss, score_real, score_fake)."""
    real = d_forward(dp, x, n_layers, alpha, arch)
    score_real = real.mean()
    with torch.no_grad():
        fake = g_forward(gp, z, n_layers, alpha, arch)                      # .detach(), :26
    score_fake = d_forward(dp, fake, n_layers, alpha, arch).mean()
    loss = -score_real + score_fake
    if drift > 0:
        loss = loss + drift * torch.square(real).mean()                     # :44-45
    return loss, score_real, score_fake


def grad_penalty(gp, dp, x, z, eps, n_layers, alpha, arch, lam=10.0, return_grad=False):
    """D_grad_pen_loss.forward (loss_functions.py:157-180); ``eps`` is the U[0,1) draw of :170."""
    with torch.no_grad():
        x_tilde = g_forward(gp, z, n_layers, alpha, arch)
    x_hat = (eps * x + (1 - eps) * x_tilde).detach().requires_grad_()
    out = d_forward(dp, x_hat, n_layers, alpha, arch)
    g = torch.autograd.grad(outputs=out.sum(), inputs=x_hat, create_graph=True)[0]
    pen = lam * torch.mean((g.norm(2, dim=(1, 2, 3)) - 1) ** 2)
    return (pen, g) if return_grad else pen


def g_w_loss(gp, dp, z, n_layers, alpha, arch):
    """G_W_loss.forward (loss_functions.py:59-74)."""
    return -d_forward(dp, g_forward(gp, z, n_layers, alpha, arch), n_layers, alpha, arch).mean()


# ----------------------------------------------------------------------------------------------
# Adam (train.py:220-225: optim.Adam(lr, betas=(beta1, 0.999)), eps 1e-8, no weight decay)
# ----------------------------------------------------------------------------------------------


class Adam:
    """torch.optim.Adam semantics restated: parameters whose grad is None are skipped and their
    step count does not advance."""

    def __init__(self, params: dict, lr=1e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.p, self.lr, self.b1, self.b2, self.eps = params, lr, beta1, beta2, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = {k: 0 for k in params}

    @torch.no_grad()
    def step(self, grads: dict):
        for k, g in grads.items():
            if g is None:
                continue
            self.t[k] += 1
            t = self.t[k]
            self.m[k].lerp_(g, 1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            self.p[k].addcdiv_(self.m[k], denom, value=-self.lr / bc1)


# ----------------------------------------------------------------------------------------------
# one training iteration (train.py:350-394), n_critic = 1
# ----------------------------------------------------------------------------------------------


class Trainer:
    """Holds G/D parameters (leaf tensors) plus the two Adam states and runs pggan_train's inner
    iteration.  RNG draws happen in the reference's order: z (D_W) -> z (GP) -> eps -> z (G); they
    can be injected for parity tests."""

    def __init__(self, arch: Arch = None, seed=1, res=None, alpha=1.0, lr=1e-4, beta1=0.5,
                 lam=10.0, drift=1e-3, dtype=torch.float32, params=None):
        self.arch = arch or Arch()
        gp, dp = params if params is not None else build_params(self.arch, seed)
        self.gp = {k: v.to(dtype).clone().requires_grad_() for k, v in gp.items()}
        self.dp = {k: v.to(dtype).clone().requires_grad_() for k, v in dp.items()}
        self.res = res or self.arch.size_init
        self.alpha = float(alpha)
        self.n_layers = n_layers_for(self.res, self.arch)
        self.lam, self.drift = lam, drift
        self.opt_g = Adam(self.gp, lr, beta1)
        self.opt_d = Adam(self.dp, lr, beta1)

    def draw(self, batch, generator=None):
        zs = [sample_latent((batch, self.arch.latent_dim), generator) for _ in range(2)]
        eps = torch.rand((batch, 1, 1, 1), generator=generator)
        zs.append(sample_latent((batch, self.arch.latent_dim), generator))
        return zs[0], zs[1], eps, zs[2]

    def d_losses(self, x, z1, z2, eps):
        a = (self.n_layers, self.alpha, self.arch)
        loss, sr, sf = d_w_loss(self.gp, self.dp, x, z1, *a, drift=self.drift)
        pen = grad_penalty(self.gp, self.dp, x, z2, eps, *a, lam=self.lam)
        return loss + pen, sr, sf, pen

    def _grads(self, loss, params, names):
        gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        return dict(zip(names, gs))

    def draw_critic(self, batch, generator=None):
        """(z, z, eps) of one critic round; with lam == 0 D_grad_pen_loss draws nothing (loss_functions.py:159)."""
        z1 = sample_latent((batch, self.arch.latent_dim), generator)
        if self.lam > 0:
            return z1, sample_latent((batch, self.arch.latent_dim), generator), torch.rand((batch, 1, 1, 1), generator=generator)
        return z1, None, None

    def iteration(self, x, draws=None, n_critic=1):
        """Returns dict(score_real, score_fake, D_loss, G_loss, D_grad_pen) as Python floats.
        n_critic = k > 1: k critic rounds on the same images with fresh draws, statistics of the last round
        (train.py:356-366); n_critic = 0: the critic losses are evaluated once for the statistics, no update
        (train.py:369-374).  `draws` = (z1, z2, eps, z3) is only accepted for the plain n_critic = 1 iteration."""
        dt = next(iter(self.gp.values())).dtype
        x = x.to(dt)
        a = (self.n_layers, self.alpha, self.arch)
        z3 = None
        for j in range(max(n_critic, 1)):
            if draws is not None:
                assert n_critic == 1
                z1, z2, eps, z3 = [t.to(dt) for t in draws]
            else:
                z1, z2, eps = [t if t is None else t.to(dt) for t in self.draw_critic(x.shape[0])]
            d_total, sr, sf = d_w_loss(self.gp, self.dp, x, z1, *a, drift=self.drift)
            if self.lam > 0:
                pen = grad_penalty(self.gp, self.dp, x, z2, eps, *a, lam=self.lam)
                d_total = d_total + pen
            else:
                pen = torch.zeros(())                                            # loss_functions.py:177-178
            if n_critic > 0:
                names = active_d_names(self.n_layers, self.alpha, self.arch)
                self.last_d_grads = self._grads(d_total, self.dp, names)
                self.opt_d.step(self.last_d_grads)                               # train.py:365-366
        if z3 is None:
            z3 = sample_latent((x.shape[0], self.arch.latent_dim)).to(dt)
        g_loss = g_w_loss(self.gp, self.dp, z3, self.n_layers, self.alpha, self.arch)
        names = active_g_names(self.n_layers, self.alpha, self.arch)
        self.last_g_grads = self._grads(g_loss, self.gp, names)
        self.opt_g.step(self.last_g_grads)                                       # train.py:384-385
        return {'score_real': sr.item(), 'score_fake': sf.item(), 'D_loss': d_total.item(),
                'G_loss': g_loss.item(), 'D_grad_pen': pen.item()}


def synthetic_images(batch, res, seed=7):
    """U[-1,1) 1-channel images from a private CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, res, res, generator=g) * 2 - 1
