"""Generator_PG / Discriminator_PG with the reference's module API (reference models.py:272-616) whose
forward / backward / double-backward run in the hand-written sm_100a kernels of libngan_b200.so.

What is kept identical to the reference, because train.py / eval.py / Checkpointer
touch it: constructor signatures, attribute names, the sub-module tree (`layers`, `conv_block_list`,
`ToIm_list`/`ToIm`, `FromIm_list`/`FromIm`) and therefore the `state_dict()` keys at every structural state,
`increase_resolution / advance_transition / set_resolution / from_state_dict`, `saved_attrs`, assertion
messages, and the RNG consumption at construction (nn.Conv2d / nn.Linear default init followed by
kaiming_normal_, models.py:31-34) so that the same seed gives the same initial weights.

What is different: the sub-modules only own parameters.  `forward` hands the whole network to
neuron_gan_b200.engine, which launches the fused kernels; gradients flow through two network-level
autograd Functions (forward, and a differentiable backward for the WGAN-GP double backward).  There is no
CPU or eager-PyTorch fallback: calling a network on a non-CUDA tensor raises.
"""
import math
import re
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

try:  # dropped into the reference tree: honour its config module for the defaults (models.py:15-18)
    from configs import config as _config
    latent_dim_default = _config.latent_dim
    image_size_default = _config.image_size
    N_colors_default = _config.N_colors
    LeakyReLU_neg_slope_default = _config.LeakyReLU_leak
except Exception:  # noqa: BLE001 - standalone use: configs/config.py:58-61 defaults
    latent_dim_default = 512
    image_size_default = 512
    N_colors_default = 1
    LeakyReLU_neg_slope_default = 0.2

from .legacy import Discriminator_dcgan, Discriminator_wgan, Generator_dcgan, Generator_wgan  # noqa: E402,F401

# the reference's `from models import *` surface (models.py:10-12)
__all__ = ['Generator_dcgan', 'Discriminator_dcgan', 'Generator_wgan', 'Discriminator_wgan', 'Generator_PG',
           'Discriminator_PG']


def _default_device():
    """Where from_state_dict(filename) puts a network when no device is given (eval.py:23): the current CUDA device;
    the CPU only when there is none (checkpoint surgery still works there, forward passes do not)."""
    return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')


def kaiming_init(model: nn.Module, neg_slope=LeakyReLU_neg_slope_default):
    """He-normal weights, zero bias (reference models.py:31-34)."""
    torch.nn.init.kaiming_normal_(model.weight, a=neg_slope, mode='fan_in', nonlinearity='leaky_relu')
    if model.bias is not None:
        model.bias.data.zero_()


# ---------------------------------------------------------------------------------------------------------
# parameter-holding sub-modules (same names / nesting as the reference so state-dict keys match)
# ---------------------------------------------------------------------------------------------------------
class _Marker(nn.Module):
    """Parameter-free placeholder that keeps nn.Sequential indices (and print(net)) like the reference's."""

    def __init__(self, text):
        super().__init__()
        self._text = text

    def extra_repr(self):
        return self._text

    def forward(self, x):
        raise RuntimeError('neuron_gan_b200 sub-modules are parameter holders; call the network, not its layers')


class Interpolate(_Marker):
    def __init__(self, scale_factor, mode='bilinear'):
        super().__init__(f'scale_factor={scale_factor}, mode={mode}')
        self.scale_factor, self.mode = scale_factor, mode


class PixelNorm(_Marker):
    def __init__(self, epsilon=1e-8):
        super().__init__(f'epsilon={epsilon}')
        self.epsilon = epsilon


class _Holder:
    """Mixin: these layers are executed by the fused engine, never on their own."""

    def forward(self, x):
        raise RuntimeError('neuron_gan_b200 sub-modules are parameter holders; call the network, not its layers')


def _equalized_scale(n_connections, act_func):
    gain = torch.nn.init.calculate_gain(nonlinearity=act_func[0], param=act_func[1]) if act_func is not None else 1
    return gain / math.sqrt(n_connections)


class Conv2d_normalized(_Holder, nn.Conv2d):
    """Equalised-LR conv (reference models.py:172-204): y = conv(weight_scale * x, W) + b."""

    def __init__(self, *args, scale_mode='fan_in', act_func=('leaky_relu', LeakyReLU_neg_slope_default), **kwargs):
        super().__init__(*args, **kwargs)
        kaiming_init(self)
        if scale_mode not in ('fan_in', 'fan_out'):
            raise ValueError('{} is not a supported mode', scale_mode)
        self.weight_scale_mode = scale_mode
        n = (self.weight.shape[1] if scale_mode == 'fan_in' else self.weight.shape[0]) * int(np.prod(self.kernel_size))
        self.scale_value = float(np.float32(_equalized_scale(n, act_func)))   # host mirror used by the kernels
        self.register_buffer('weight_scale', torch.tensor(self.scale_value), persistent=False)


class Linear_normalized(_Holder, nn.Linear):
    """Equalised-LR linear (reference models.py:208-241)."""

    def __init__(self, *args, scale_mode='fan_in', act_func=('leaky_relu', LeakyReLU_neg_slope_default), **kwargs):
        super().__init__(*args, **kwargs)
        kaiming_init(self)
        if scale_mode not in ('fan_in', 'fan_out'):
            raise ValueError('{} is not a supported mode', scale_mode)
        self.weight_scale_mode = scale_mode
        n = self.weight.shape[1] if scale_mode == 'fan_in' else self.weight.shape[0]
        self.scale_value = float(np.float32(_equalized_scale(n, act_func)))
        self.register_buffer('weight_scale', torch.tensor(self.scale_value), persistent=False)


class ToImage(_Holder, nn.Module):
    """1x1 conv to colour space + tanh (reference models.py:133-149); key `layers.0.weight`."""

    def __init__(self, in_channels, N_colors):
        super().__init__()
        self.in_channels, self.N_colors = in_channels, N_colors
        conv = nn.Conv2d(in_channels, N_colors, kernel_size=1, stride=1, padding=0, bias=False)
        kaiming_init(conv)
        self.layers = nn.Sequential(conv, nn.Tanh())

    @property
    def weight(self):
        return self.layers[0].weight

    def extra_repr(self):
        return 'in_channels={}, N_colors={}'.format(self.in_channels, self.N_colors)


class FromImage(_Holder, nn.Module):
    """1x1 conv from colour space, with bias, no activation (reference models.py:156-165); keys `conv.*`."""

    def __init__(self, N_colors, out_channels):
        super().__init__()
        self.out_channels, self.N_colors = out_channels, N_colors
        self.conv = nn.Conv2d(N_colors, out_channels, kernel_size=1, stride=1, padding=0)
        kaiming_init(self.conv)

    def extra_repr(self):
        return 'N_colors={}, out_channels={}'.format(self.N_colors, self.out_channels)


class Conv2d_scale_block(_Holder, nn.Sequential):
    """resample -> conv/lrelu/PN -> conv/lrelu/PN (reference models.py:245-268); convs at indices 1 and 4."""

    def __init__(self, in_channels, out_channels, kernel_size, padding=1, scale_factor=None,
                 LeakyReLU_neg_slope=LeakyReLU_neg_slope_default):
        super().__init__()
        if scale_factor < 1:
            self.append(nn.AvgPool2d(kernel_size=int(1 / scale_factor)))
        else:
            self.append(Interpolate(scale_factor=scale_factor, mode='bilinear'))
        act = ('leaky_relu', LeakyReLU_neg_slope)
        self.append(Conv2d_normalized(in_channels, out_channels, kernel_size, stride=1, padding=padding,
                                      padding_mode='zeros', bias=False, act_func=act))
        self.append(nn.LeakyReLU(negative_slope=LeakyReLU_neg_slope))
        self.append(PixelNorm())
        self.append(Conv2d_normalized(out_channels, out_channels, kernel_size, stride=1, padding=padding,
                                      padding_mode='zeros', bias=False, act_func=act))
        self.append(nn.LeakyReLU(negative_slope=LeakyReLU_neg_slope))
        self.append(PixelNorm())

    @property
    def conv1(self):
        return self[1]

    @property
    def conv2(self):
        return self[4]


# ---------------------------------------------------------------------------------------------------------
# shared progressive-growing bookkeeping
# ---------------------------------------------------------------------------------------------------------
class _ProgressiveNet(nn.Module):
    alpha: torch.Tensor

    def _init_progressive(self, N_features_per_layer, image_size_init, LeakyReLU_neg_slope, N_colors):
        N_scaling = len(N_features_per_layer) - 1
        self.N_features_per_layer = N_features_per_layer
        self.N_layers = 1
        self.N_layers_max = len(self.N_features_per_layer)
        self.N_colors = N_colors
        self.image_size_init = image_size_init
        self.image_size = image_size_init
        self.image_size_max = 2 ** N_scaling * image_size_init
        self.LeakyReLU_neg_slope = LeakyReLU_neg_slope

    def __setattr__(self, name, value):
        if name == 'alpha':
            object.__setattr__(self, '_alpha_key', None)       # host mirror is stale
        super().__setattr__(name, value)

    def _finish_saved_attrs(self, not_saved):
        # same construction as the reference (models.py:338-342): public instance attributes minus modules
        attrs = set(self.__dict__.keys()) - set(nn.Module.__dict__.keys())
        attrs = sorted(a for a in attrs if not a.startswith('_') and a not in not_saved)
        attrs.append('alpha')
        self.saved_attrs = attrs

    # -- host mirror of alpha: the reference branches on the device tensor every forward (a host sync,
    #    models.py:345, 517); here the value is re-read only when the buffer object or its version changed.
    def alpha_value(self) -> float:
        a = self.alpha
        key = (id(a), a._version)
        if getattr(self, '_alpha_key', None) != key:
            object.__setattr__(self, '_alpha_host', float(a.item()))
            object.__setattr__(self, '_alpha_key', key)
        return self._alpha_host

    def increase_resolution(self):
        assert self.alpha >= 1, 'The previous transition has not ended.'
        self.alpha = 0 * self.alpha
        self.N_layers += 1
        self.image_size *= 2
        assert self.image_size <= self.image_size_max, (
            f'The image size ({self.image_size}) is greater than the maximum ({self.image_size_max})')

    def set_resolution(self, res: int, alpha=1.0):
        assert res % self.image_size == 0, 'The resolution must be divisible by {}'.format(self.image_size)
        assert math.log2(res / self.image_size).is_integer(), (
            f'{res} cannot be attained by multiplying the initial resolution ({self.image_size}) by a power of 2.')
        assert res <= self.image_size_max, 'The resolution must be smaller than {}'.format(self.image_size_max)
        while self.image_size < res:
            self.increase_resolution()
            self.advance_transition(alpha if self.image_size == res else 1.0)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        from . import engine
        engine.invalidate(self)         # copy_() bumps the versions already; explicit so that .data writes are covered
        return out

    def _check_input(self, x, what):
        if not x.is_cuda:
            raise RuntimeError(f'neuron_gan_b200.{type(self).__name__} runs on CUDA (sm_100a) only; got a {x.device} '
                               f'{what}. There is no CPU fallback: move the network and its inputs to the GPU.')
        if next(self.parameters()).device != x.device:
            raise RuntimeError('network parameters and input are on different devices')


def _legacy_state_dict_surgery(state_dict, list_name, n_present, n_expected, extra_prefixes, from_start):
    """Old checkpoints kept already-merged modules in the lists (reference models.py:38-63, 411-436): drop the
    surplus entries and renumber, exactly like pop_state_dict_modules."""
    def pop(sd, prefix, n_delete):
        keys = [k for k in sd if k.startswith(prefix)]
        if not keys:
            return sd
        idx = [int(re.search(r'\d', k)[0]) for k in keys]
        n_max = max(idx) + 1
        if n_delete == 'all':
            n_delete = n_max
        assert n_delete <= n_max, 'Cannot remove more than {} layers'.format(n_max)
        removed = set(range(n_delete)) if from_start else set(range(n_max - n_delete, n_max))
        out = OrderedDict()
        for k, v in sd.items():
            if k in keys:
                i = int(re.search(r'\d', k)[0])
                if i in removed:
                    continue
                if from_start:
                    m = re.search(r'\d', k)
                    k = k[:m.start()] + str(i - n_delete) + k[m.end():]
            out[k] = v
        return out

    for name, present, expected in zip(list_name, n_present, n_expected):
        state_dict = pop(state_dict, name, present - expected)
    for prefix in extra_prefixes:
        state_dict = pop(state_dict, prefix, 'all')
    return state_dict


def _count_modules(state_dict, list_name):
    patt = re.compile(r'(?<=' + list_name + r'.)\d')
    idx = [int(patt.findall(k)[0]) for k in state_dict if patt.search(k)]
    return max(idx) + 1 if idx else 0


# ---------------------------------------------------------------------------------------------------------
# Generator
# ---------------------------------------------------------------------------------------------------------
class Generator_PG(_ProgressiveNet):
    """Progressively growing generator (reference models.py:272-444)."""

    def __init__(self, N_features_per_layer: list, image_size_init=4, latent_dim=latent_dim_default,
                 LeakyReLU_neg_slope=LeakyReLU_neg_slope_default, N_colors=N_colors_default):
        super().__init__()
        self.latent_dim = latent_dim
        self._init_progressive(N_features_per_layer, image_size_init, LeakyReLU_neg_slope, N_colors)
        self.register_buffer('alpha', torch.tensor(1.0), persistent=False)
        f = self.N_features_per_layer
        act = ('leaky_relu', LeakyReLU_neg_slope)

        self.layers = nn.Sequential()
        self.layers.append(Linear_normalized(latent_dim, f[0] * image_size_init ** 2, bias=False, act_func=act))
        self.layers.append(nn.Unflatten(dim=1, unflattened_size=(f[0], image_size_init, image_size_init)))
        self.layers.append(nn.LeakyReLU(negative_slope=LeakyReLU_neg_slope))
        self.layers.append(PixelNorm())
        self.layers.append(Conv2d_normalized(f[0], f[0], kernel_size=3, stride=1, padding=1, padding_mode='zeros',
                                             bias=False, act_func=act))
        self.layers.append(nn.LeakyReLU(negative_slope=LeakyReLU_neg_slope))
        self.layers.append(PixelNorm())

        self.conv_block_list = nn.ModuleList(
            Conv2d_scale_block(in_channels=f[i], out_channels=f[i + 1], scale_factor=2, kernel_size=3)
            for i in range(len(f) - 1))
        self.ToIm_list = nn.ModuleList(ToImage(f[i], N_colors) for i in range(len(f)))
        self.ToIm = self.ToIm_list.pop(0)
        self.upsample = Interpolate(scale_factor=2, mode='bilinear')
        self._finish_saved_attrs(['layers', 'ToIm_list', 'ToIm', 'conv_block_list'])

    def advance_transition(self, alpha_step=0.1):
        self.alpha += alpha_step
        if self.alpha >= 1.0:
            self.layers.append(self.conv_block_list.pop(0))
            self.ToIm = self.ToIm_list.pop(0)

    # ---- structure seen by the engine
    def trunk_blocks(self):
        return list(self.layers)[7:]

    def active_parameters(self):
        """Parameters that receive a gradient at the current structural state, in a fixed order."""
        ps = [self.layers[0].weight, self.layers[4].weight]
        for blk in self.trunk_blocks():
            ps += [blk.conv1.weight, blk.conv2.weight]
        ps.append(self.ToIm.weight)
        if self.alpha_value() < 1:
            blk = self.conv_block_list[0]
            ps += [blk.conv1.weight, blk.conv2.weight, self.ToIm_list[0].weight]
        return ps

    def forward(self, x):
        self._check_input(x, 'latent batch')
        if self.N_colors != 1:
            raise NotImplementedError('only N_colors == 1 is built (the reference dataset is 1-channel)')
        from . import autograd_fns
        return autograd_fns.generator_forward(self, x)

    @classmethod
    def from_state_dict(cls, filename, device=None, verbose=True):
        """Rebuild a generator from a reference-format .pth (reference models.py:394-444).  device=None (the way
        eval.py:23 calls it): the current CUDA device -- the reference defaults to the CPU, where these networks
        cannot run."""
        device = _default_device() if device is None else torch.device(device)
        saved = torch.load(filename, map_location=device, weights_only=False)
        attrs = saved['Generator_attrs']
        ctor = {k: v for k, v in attrs.items() if k in ('N_features_per_layer', 'image_size_init',
                                                         'LeakyReLU_neg_slope', 'N_colors')}
        obj = cls(**ctor)
        attrs = {k: v for k, v in attrs.items() if k in obj.saved_attrs}
        obj.set_resolution(attrs['image_size'], float(attrs['alpha']))
        state = saved['Generator_state']
        n_toim = _count_modules(state, 'ToIm_list')
        n_blk = _count_modules(state, 'conv_block_list')
        if n_toim > len(obj.ToIm_list):
            if verbose:
                print('Warning! Loaded state dict in old format. Keys will be removed to match the new format.')
            state = _legacy_state_dict_surgery(state, ['ToIm_list', 'conv_block_list'], [n_toim, n_blk],
                                               [len(obj.ToIm_list), len(obj.conv_block_list)],
                                               ['ToIm_prev', 'last_conv_block'], from_start=True)
        obj.load_state_dict(state)
        obj.to(device)                  # the reference builds on the CPU and leaves it there (eval.py runs on CPU);
        if verbose:                     # these networks only run on CUDA, so `device` is where the result lives
            print('Loaded training state from {}'.format(filename))
        return obj


# ---------------------------------------------------------------------------------------------------------
# Discriminator
# ---------------------------------------------------------------------------------------------------------
class Discriminator_PG(_ProgressiveNet):
    """Progressively growing critic (reference models.py:448-616)."""

    def __init__(self, N_features_per_layer: list, image_size_init=4, LeakyReLU_neg_slope=LeakyReLU_neg_slope_default,
                 N_colors=N_colors_default):
        super().__init__()
        self._init_progressive(N_features_per_layer, image_size_init, LeakyReLU_neg_slope, N_colors)
        self.register_buffer('alpha', torch.tensor(1.0), persistent=True)
        f = self.N_features_per_layer
        act = ('leaky_relu', LeakyReLU_neg_slope)

        self.layers = nn.Sequential()
        self.layers.append(Conv2d_normalized(f[-1], f[-1], kernel_size=3, stride=1, padding=1, padding_mode='zeros',
                                             act_func=act))
        self.layers.append(nn.LeakyReLU(negative_slope=LeakyReLU_neg_slope))
        self.layers.append(PixelNorm())
        self.layers.append(Conv2d_normalized(f[-1], 1, (image_size_init, image_size_init), stride=1, padding=0,
                                             act_func=act))
        self.layers.append(nn.Flatten())

        self.conv_block_list = nn.ModuleList(
            Conv2d_scale_block(in_channels=f[i], out_channels=f[i + 1], scale_factor=0.5, kernel_size=3)
            for i in range(len(f) - 1))
        self.FromIm_list = nn.ModuleList(FromImage(N_colors, f[i]) for i in range(len(f)))
        self.FromIm = self.FromIm_list.pop(-1)
        self.downsample = Interpolate(scale_factor=0.5, mode='bilinear')
        self._finish_saved_attrs(['layers', 'FromIm_list', 'FromIm', 'conv_block_list'])

    def advance_transition(self, alpha_step=0.1):
        self.alpha += alpha_step
        if self.alpha >= 1.0:
            self.layers.insert(0, self.conv_block_list.pop(-1))
            self.FromIm = self.FromIm_list.pop(-1)

    # ---- structure seen by the engine
    def trunk_blocks(self):
        """Blocks already merged into `layers`, highest resolution first."""
        return [m for m in self.layers if isinstance(m, Conv2d_scale_block)]

    def last_conv(self):
        return self.layers[len(self.trunk_blocks())]

    def head_conv(self):
        return self.layers[len(self.trunk_blocks()) + 3]

    def active_parameters(self):
        ps = []
        if self.alpha_value() < 1:
            blk = self.conv_block_list[-1]
            new = self.FromIm_list[-1].conv
            ps += [new.weight, new.bias, blk.conv1.weight, blk.conv2.weight]
        ps += [self.FromIm.conv.weight, self.FromIm.conv.bias]
        for blk in self.trunk_blocks():
            ps += [blk.conv1.weight, blk.conv2.weight]
        last, head = self.last_conv(), self.head_conv()
        ps += [last.weight, last.bias, head.weight, head.bias]
        return ps

    def forward(self, x):
        self._check_input(x, 'image batch')
        if self.N_colors != 1:
            raise NotImplementedError('only N_colors == 1 is built (the reference dataset is 1-channel)')
        from . import autograd_fns
        return autograd_fns.discriminator_forward(self, x)

    @classmethod
    def from_state_dict(cls, filename, device=None, verbose=True):
        """Rebuild a critic from a reference-format .pth (reference models.py:566-616); device=None: current CUDA device."""
        device = _default_device() if device is None else torch.device(device)
        saved = torch.load(filename, map_location=device, weights_only=False)
        attrs = saved['Discriminator_attrs']
        ctor = {k: v for k, v in attrs.items() if k in ('N_features_per_layer', 'image_size_init',
                                                         'LeakyReLU_neg_slope', 'N_colors')}
        obj = cls(**ctor)
        attrs = {k: v for k, v in attrs.items() if k in obj.saved_attrs}
        obj.set_resolution(attrs['image_size'], float(attrs['alpha']))
        state = saved['Discriminator_state']
        n_from = _count_modules(state, 'FromIm_list')
        n_blk = _count_modules(state, 'conv_block_list')
        if n_from > len(obj.FromIm_list):
            if verbose:
                print('Warning! Loaded state dict in old format. Keys will be removed to match the new format.')
            state = _legacy_state_dict_surgery(state, ['FromIm_list', 'conv_block_list'], [n_from, n_blk],
                                               [len(obj.FromIm_list), len(obj.conv_block_list)],
                                               ['FromIm_prev', 'first_conv_block'], from_start=False)
        obj.load_state_dict(state)
        obj.to(device)
        if verbose:
            print('Loaded training state from {}'.format(filename))
        return obj
