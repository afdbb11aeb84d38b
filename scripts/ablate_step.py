"""What is each kernel family worth in the captured iteration?  Replaces one family at a time by a no-op (results
are then wrong -- this only measures time) and reports the graph-replayed ms / iteration.
usage (B200): python scripts/ablate_step.py [res] [alpha] [batch]"""
import sys

import torch

sys.path.insert(0, '.')
from neuron_gan_b200 import engine, ops
from neuron_gan_b200.train_step import TrainStep, build_networks

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16


def measure(tag, patch=None, undo=None, **kw):
    if patch:
        patch()
    G, D = build_networks(res, alpha, seed=1, device='cuda')
    step = TrainStep(G, D)
    for k, v in kw.items():
        setattr(step, k, v)
    g = torch.Generator().manual_seed(3)
    xs = [(torch.rand(B, 1, res, res, generator=g) * 2 - 1).cuda() for _ in range(4)]
    draws = [step.draw(B, 'cuda') for _ in range(4)]
    for i in range(4):
        step(xs[i % 4], draws[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for i in range(20):
            step(xs[i % 4], draws[i % 4])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(f'{tag:34s} {best:7.3f} ms  ({step.launches_per_step} launches, {step.last_run})', flush=True)
    if undo:
        undo()
    return best


saved = {}


def noop(name, ret=None):
    def patch():
        saved[name] = getattr(ops, name)
        setattr(ops, name, (lambda *a, **k: ret(*a, **k)) if callable(ret) else (lambda *a, **k: ret))
    def undo():
        setattr(ops, name, saved[name])
    return patch, undo


base = measure('baseline')
measure('three generator passes as one batch', merge_g_forward=True)
measure('generator passes separate', merge_g_forward=False)
measure('four graphs (segment_graphs)', segment_graphs=True)
for ns in (2, 4, 6):
    def setn(n=ns):
        engine._Side.n_streams = n; engine._Side.streams = []
    def unsetn():
        engine._Side.n_streams = 6; engine._Side.streams = []
    measure(f'{ns} wgrad side streams', setn, unsetn)
measure('factored linear grad + plain Adam', factor_linear=True)
measure('factored linear grad fused into Adam', factor_linear=True, fuse_linear_adam=True)
p, u = noop('conv3x3_wgrad')
measure('no conv3x3_wgrad', p, u)
side = engine._Side.enabled
def off(): engine._Side.enabled = False
def on(): engine._Side.enabled = side
measure('wgrad inline (no side streams)', off, on)
measure('no chain fork', fork_chains=False)
_cache = {}
def fake_up(x):
    k = tuple(x.shape)
    if k not in _cache:
        B_, C8, H, W, e = x.shape
        _cache[k] = torch.zeros((B_, C8, 2 * H, 2 * W, e), dtype=x.dtype, device=x.device)
    return _cache[k]
p, u = noop('upsample2x', fake_up)
measure('no upsample2x', p, u)
def fake_pool(x):
    k = ('p',) + tuple(x.shape)
    if k not in _cache:
        B_, C8, H, W, e = x.shape
        _cache[k] = torch.zeros((B_, C8, H // 2, W // 2, e), dtype=x.dtype, device=x.device)
    return _cache[k]
p, u = noop('avgpool2', fake_pool)
measure('no avgpool2', p, u)
def fake_adam(*a, **k): return None
p, u = noop('adam_multi')
measure('no adam', p, u)
