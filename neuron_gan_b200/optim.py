"""FusedAdam: torch.optim.Adam semantics (reference train.py:220-225) with the whole step as one multi-tensor
kernel launch per network.  Keeps torch's state layout (`step`, `exp_avg`, `exp_avg_sq` per parameter) and
`param_groups[...]['lr']`, which the reference's update_lr (train.py:250-265) mutates."""
import math

import torch

from . import engine, ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            lr = group['lr']
            entries, touched = [], []
            for p in group['params']:
                if p.grad is None:            # inactive blocks: skipped, step count does not advance
                    continue
                if not p.is_cuda:
                    raise RuntimeError('FusedAdam runs on CUDA parameters only (no CPU fallback)')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                t = st['step']
                shadow = None
                if p.dim() == 2:              # the generator's Linear weight keeps a same-layout bf16 shadow
                    ent = engine._cache_get(p)
                    if ent is not None and 'shadow' in ent:
                        shadow = ent['shadow']
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                entries.append(dict(p=p, g=g, m=st['exp_avg'], v=st['exp_avg_sq'], shadow=shadow,
                                    step_size=lr / (1 - beta1 ** t), inv_bc2_sqrt=1 / math.sqrt(1 - beta2 ** t)))
                touched.append((p, shadow is not None))
            for i in range(0, len(entries), 64):
                ops.adam_multi(entries[i:i + 64], beta1, beta2, group['eps'])
            for p, fresh in touched:
                engine.mark_updated(p, shadow_is_fresh=fresh)
        return loss
