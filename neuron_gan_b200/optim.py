"""FusedAdam: torch.optim.Adam semantics (reference train.py:220-225) with the whole step as one multi-tensor
kernel launch per network.  Keeps torch's state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) and
`param_groups[...]['lr']`, which the reference's update_lr (train.py:250-265) mutates.

`step()` = `advance()` + `launch()`.  With `capturable=True` the two per-parameter scalars that change every step
(lr / (1 - beta1^t) and 1 / sqrt(1 - beta2^t)) live in a small device buffer that `advance()` refreshes from the
host, so a `launch()` captured in a CUDA graph stays valid across steps and learning-rate changes."""
import math

import torch

from . import engine, ops

_CHUNK = 48          # tensors per launch (kAdamMaxTensors in csrc/adam.cu)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.capturable = capturable
        self._dyn = None            # device [n_params, 2] fp32, capturable mode
        self._ring = None           # pinned staging for the per-step scalars
        self._slot = {}

    # -- helpers -----------------------------------------------------------------------------------------
    def _active(self):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:            # inactive blocks: skipped, step count does not advance
                    continue
                if not p.is_cuda:
                    raise RuntimeError('FusedAdam runs on CUDA parameters only (no CPU fallback)')
                yield group, p

    def _state(self, p):
        st = self.state[p]
        if not st:
            st['step'] = 0
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _ensure_dyn(self, device):
        if self._dyn is None:
            n = 0
            for group in self.param_groups:
                for p in group['params']:
                    self._slot[id(p)] = n
                    n += 1
            self._dyn = torch.zeros((n, 2), dtype=torch.float32, device=device)
        return self._dyn

    @staticmethod
    def _scalars(group, t):
        beta1, beta2 = group['betas']
        return group['lr'] / (1 - beta1 ** t), 1 / math.sqrt(1 - beta2 ** t)

    # -- the two halves of a step ----------------------------------------------------------------------------
    @torch.no_grad()
    def advance(self):
        """Host half: bump every active parameter's step count; in capturable mode also ship this step's
        bias-correction scalars to the device buffer (one small async copy from pinned memory)."""
        host, dev, k = None, None, None
        for group, p in self._active():
            st = self._state(p)
            st['step'] += 1
            if self.capturable:
                if host is None:
                    dev = self._ensure_dyn(p.device)
                    if self._ring is None:
                        from .utils import PinnedRing
                        self._ring = PinnedRing([tuple(dev.shape)])
                        for (buf,) in self._ring.sets:
                            buf.zero_()
                    k, (host,) = self._ring.acquire()
                    rows = host.numpy()
                a, b = self._scalars(group, st['step'])
                i = self._slot[id(p)]
                rows[i, 0] = a
                rows[i, 1] = b
        if host is not None:
            dev.copy_(host, non_blocking=True)
            self._ring.release(k)

    @torch.no_grad()
    def launch(self, factored=None):
        """Device half: one multi-tensor kernel per <= 48 tensors, on the current stream (capturable).
        factored: optional {id(param): dict(ga, z, gscale, K, C, S, b_per_seg, n_seg, ga_seg_stride, z_seg_stride,
        g_out)} -- the generator's Linear weight, whose gradient is handed over as its two factors and formed inside
        the Adam pass (ops.adam_linear_factored) instead of being read from p.grad."""
        factored = factored or {}
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            entries, touched = [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self._state(p)
                if id(p) in factored:
                    f = factored[id(p)]
                    ent = engine._cache_get(p)
                    shadow = ent['shadow'] if ent is not None and 'shadow' in ent and ent['shadow'].device == p.device else None
                    if self.capturable:
                        self._ensure_dyn(p.device)
                        a, b, dyn = 0.0, 0.0, self._dyn[self._slot[id(p)]]
                    else:
                        (a, b), dyn = self._scalars(group, max(st['step'], 1)), None
                    ops.adam_linear_factored(p, st['exp_avg'], st['exp_avg_sq'], shadow, f['ga'], f['z'], f['K'], f['C'],
                                             f['S'], f['gscale'], a, b, dyn, beta1, beta2, group['eps'],
                                             b_per_seg=f.get('b_per_seg'), ga_seg_stride=f.get('ga_seg_stride', 0),
                                             z_seg_stride=f.get('z_seg_stride', 0), n_seg=f.get('n_seg', 1),
                                             g_out=f.get('g_out'))
                    touched.append((p, shadow is not None))
                    continue
                # bf16 operand images of the GEMM weights (linear stem, 3x3 convs) are rewritten in the same pass
                shadow, shadow_dims, shadow_kind = None, None, 0
                ent = engine._cache_get(p) if p.dim() in (2, 4) else None
                if ent is not None and 'shadow' in ent and ent['shadow'].device == p.device:
                    shadow, shadow_dims, shadow_kind = ent['shadow'], ent['shadow_dims'], ent['shadow_kind']
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if self.capturable:
                    self._ensure_dyn(p.device)
                    a, b, dyn = 0.0, 0.0, self._dyn[self._slot[id(p)]]
                else:
                    (a, b), dyn = self._scalars(group, max(st['step'], 1)), None
                entries.append(dict(p=p, g=g, m=st['exp_avg'], v=st['exp_avg_sq'], shadow=shadow, shadow_dims=shadow_dims, shadow_kind=shadow_kind,
                                    step_size=a, inv_bc2_sqrt=b, dyn=dyn))
                touched.append((p, shadow is not None))
            for i in range(0, len(entries), _CHUNK):
                ops.adam_multi(entries[i:i + _CHUNK], beta1, beta2, group['eps'])
            for p, fresh in touched:
                engine.mark_updated(p, shadow_is_fresh=fresh)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.advance()
        self.launch()
        return loss
