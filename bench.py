#!/usr/bin/env python
"""Benchmark of the north-star path: full D+G WGAN-GP training iterations (reference train.py:350-394) of the
progressive-growing GAN, images/sec.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's own step on the host CPU cores
    python bench.py --workload eval                           # generator-only inference (BASELINE config 5)

One "step" = one iteration: critic step (D_W loss + gradient penalty, backward incl. double backward, Adam)
followed by the generator step (G_W loss, backward, Adam) on one batch of synthetic 1-channel images.
Default workload: the 512x512 final phase at 16 images per GPU (BASELINE.json config 3 is 16/GPU x 8 GPUs);
weak scaling keeps 16 images per GPU.  `--res/--alpha/--batch` select other phases for sweeps.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events over K steps with inputs resident in HBM;
`e2e` is the same K steps through the public TrainStep call with the images copied from pinned host memory,
the latent / epsilon draws made on the CPU generator and copied, and the statistics read back, every step.
Both are the MEDIAN of `--rounds` timed rounds of K steps (`rounds` in the line lists every round).

`roofline` is the STEP-level figure: the layer-granular minimum HBM bytes of one iteration (SURVEY.md section 8d,
DESIGN.md section 5) / the measured iteration time / the measured copy bandwidth.  `roofline.family` and
`roofline.wgrad_family` weight every launch shape of the tcgen05 conv kernels / the weight-gradient kernel by its
launches per step, each shape graph-timed in isolation; `roofline.best_launch` is the single best launch.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train images/sec (D+G WGAN-GP step)'
EVAL_METRIC = 'eval images/sec (generator-only inference, z -> image -> host)'

# SURVEY.md section 8d (measured from the reference modules with hooks, alpha = 1 rows): per image and iteration
MIN_BYTES_MB = {16: 5.26, 32: 13.3, 64: 29.5, 128: 82.4, 256: 212.8, 512: 640.1}        # 15*B_D + 5*B_G, bf16
NECESSARY_GFLOP = {16: 1.571, 32: 4.290, 64: 7.011, 128: 12.159, 256: 23.051, 512: 43.706}   # 14*F_D + 5*F_G - lin
EVAL_BYTES_MB_512 = 59.3      # B_G per image, per-layer, bf16 (eval config)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--rounds', type=int, default=3, help='timed rounds of --steps; the median is reported')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='train', choices=['train', 'eval'])
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--alpha', type=float, default=1.0)
    ap.add_argument('--batch', type=int, default=16, help='images per GPU')
    ap.add_argument('--cpu-batch', type=int, default=0, help='batch of the CPU baseline sample (0 = auto)')
    ap.add_argument('--eval-n', default='20,64,256,1024,4096', help='sample counts of the eval workload')
    ap.add_argument('--eval-dtype', default='bf16', choices=['bf16', 'f32'], help='image dtype of the eval workload')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-eager-baseline', action='store_true', help='skip the eager-PyTorch-on-GPU baseline')
    ap.add_argument('--no-graph', action='store_true', help='launch kernel by kernel instead of replaying a CUDA graph')
    ap.add_argument('--no-profile', action='store_true', help='skip the per-kernel timing passes (roofline families)')
    ap.add_argument('--dump-kernels', default='', help='write the full per-kernel table of the timing pass here')
    return ap.parse_args()


def workload_name(res, alpha):
    phase = 'stable' if alpha >= 1 else f'fade-in(alpha={alpha:g})'
    return f'pggan_{res}x{res}_{phase}_wgan-gp_step'


def synthetic_images(batch, res, seed=7):
    """U[-1,1) 1-channel images from a private CPU generator."""
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, res, res, generator=g) * 2 - 1


# ------------------------------------------------------------------------------------------------ CPU / reference arms
def cpu_batch_for(res):
    return {16: 16, 32: 16, 64: 16, 128: 8, 256: 4, 512: 2}.get(res, 2)


def time_reference_step(res, alpha, batch, steps, warmup, device='cpu'):
    """The reference's own step.  kind 'reference': the UNMODIFIED reference modules (oracle/_ref, copied by
    oracle/make_ref.sh, or /root/reference); kind 'port': the oracle (CPU restatement) when no copy is present.
    Returns (images/s, ms/step, threads, kind)."""
    import torch
    from oracle import ref_loader
    if device == 'cpu':
        torch.set_num_threads(os.cpu_count() or 1)
    else:
        torch.backends.cudnn.benchmark = True          # the reference's only CUDA knob (train.py:128-144)
    if ref_loader.available():
        ips, ms = ref_loader.timed_iterations(res, alpha, batch, steps, warmup, device=device)
        return ips, ms, torch.get_num_threads(), 'reference'
    if device != 'cpu':
        raise RuntimeError('the GPU eager baseline needs the reference copy (bash oracle/make_ref.sh)')
    from oracle import pggan_oracle as O
    tr = O.Trainer(O.Arch(), seed=1, res=res, alpha=alpha)
    x = O.synthetic_images(batch, res)
    for _ in range(warmup):
        tr.iteration(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(x)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, torch.get_num_threads(), 'port'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    if args.workload == 'eval':
        return run_reference_eval(args)
    batch = args.cpu_batch or cpu_batch_for(args.res)
    # bound the sample so the whole run stays within minutes on the box's host cores
    steps, warmup = min(args.steps, 5), max(min(args.warmup, 2), 1)
    ips, ms, cores, kind = time_reference_step(args.res, args.alpha, batch, steps, warmup)
    what = 'the unmodified reference modules' if kind == 'reference' else 'the oracle port of the reference'
    sample = f'{steps} timed + {warmup} warm-up iterations of the same step at batch {batch} ({what}, fp32, CPU)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.res, args.alpha), 'resolution': args.res, 'alpha': args.alpha,
                       'batch': batch, 'device': 'cpu'},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def run_reference_eval(args):
    """The reference's gen_samples (utils.py:346-355) on the host cores, a bounded sample of n = 8 images."""
    import torch
    from oracle import ref_loader
    torch.set_num_threads(os.cpu_count() or 1)
    n = 8
    if ref_loader.available():
        _, _, ref_utils = ref_loader.load()
        G, _ = ref_loader.build_nets(args.res, 1.0)
        kind = 'reference'
        fn = lambda: ref_utils.gen_samples(G, N_images=n, seed=0)[0]
    else:
        from oracle import pggan_oracle as O
        arch = O.Arch()
        gp, _ = O.build_params(arch, 1)
        z = O.sample_latent((n, 512))
        kind = 'port'
        with torch.no_grad():
            fn = lambda: O.g_forward(gp, z, O.n_layers_for(args.res, arch), 1.0, arch)
    with torch.no_grad():
        fn()
        t0 = time.perf_counter()
        for _ in range(2):
            fn()
        dt = (time.perf_counter() - t0) / 2
    ips = n / dt
    line = {'impl': 'reference', 'metric': EVAL_METRIC, 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': 2, 'warmup': 1, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'pggan_{args.res}x{args.res}_generator_inference', 'n_images': n, 'device': 'cpu'},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': torch.get_num_threads(), 'kind': kind,
                             'sample': f'2 timed + 1 warm-up calls of gen_samples(N_images={n}) (fp32, CPU)'},
            'e2e': {'value': ips, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '20', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [f.strip() for f in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [f for t, f in self.samples if t0 <= t <= t1 and len(f) >= 7] or [f for _, f in self.samples if len(f) >= 7]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        sm = [float(r[0]) for r in rows if r[0].replace('.', '').isdigit()]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(rows[0][1]),
                'power_w_max': max(float(r[2]) for r in rows), 'samples': len(rows), 'reasons': reasons}


CONV_NAMES = ('ngan_conv3x3_fwd', 'ngan_conv3x3_fwd_toim', 'ngan_conv3x3_dgrad', 'ngan_conv3x3_dgrad_pn',
              'ngan_conv3x3_dbl')


def algorithmic_bytes(name, a):
    """Minimum HBM bytes one launch must move (inputs + outputs once, bf16 features, fp32 scales/images);
    DESIGN.md section 5 states the same per-pixel figures."""
    if name == 'ngan_conv3x3_fwd':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 2 * co + 4)
    if name == 'ngan_conv3x3_fwd_toim':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 4)               # (+ 2*co + 4 when y / r are stored as well: not counted)
    if name == 'ngan_conv3x3_dgrad':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 2 * co)
    if name == 'ngan_conv3x3_dgrad_pn':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * co + 2 * ci + 4 + 2 * ci)
    if name == 'ngan_conv3x3_dbl':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 4 * co + 4 + 4 * co)
    if name == 'ngan_conv3x3_wgrad':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 2 * co)
    if name in ('ngan_upsample2x',):
        B, C, H, W = a[-4:]
        return B * H * W * C * 2 * 5
    if name in ('ngan_avgpool2',):
        B, C, H, W = a[-4:]
        return B * H * W * C * 2 * 1.25
    if name == 'ngan_pn_bwd':
        B, C, H, W = a[-4:]
        return B * H * W * (C * 2 * 2.25 + 4)
    if name == 'ngan_up2_bwd_pn_bwd':
        B, C, H, W = a[-4:]
        return B * H * W * (C * 2 * 6 + 4)
    return None


def conv_flops(name, a):
    if name.startswith('ngan_conv3x3'):
        B, ci, co, H, W = a[-5:]
        return 2.0 * 9 * ci * co * B * H * W
    return None


def time_launch_isolated(name, dims, iters=20):
    """One launch shape of the conv / wgrad kernels, graph-timed on its own: `iters` launches over rotating buffers
    (working set > L2 for the large shapes) captured into one CUDA graph; best of 3 replays.  Microseconds."""
    import torch
    from neuron_gan_b200 import ops as o
    B, cin, cout, H, W = dims
    dev = 'cuda'
    per = B * H * W * (cin + cout) * 2
    n_buf = max(2, min(8, int(math.ceil(160e6 / max(per, 1)))))
    mk = lambda C: [torch.randn(B, C // 8, H, W, 8, device=dev).bfloat16() for _ in range(n_buf)]
    xin, xout = mk(cin), mk(cout)
    w = torch.randn(cout, cin, 3, 3, device=dev) / math.sqrt(cin * 9)
    w_fwd, w_dg = o.prep_conv_weight(w)
    r = torch.rand(B, H, W, device=dev) + 0.5
    s, leak = 0.1, 0.2
    if name == 'ngan_conv3x3_fwd':
        fn = lambda i: o.conv3x3_fwd(xin[i % n_buf], w_fwd, None, s, leak, cout)
    elif name == 'ngan_conv3x3_fwd_toim':
        tw = torch.randn(cout, device=dev)
        fn = lambda i: o.conv3x3_fwd_toim(xin[i % n_buf], w_fwd, None, s, leak, cout, tw, want_y=False, want_r=False)
    elif name == 'ngan_conv3x3_dgrad':
        fn = lambda i: o.conv3x3_dgrad(xout[i % n_buf], w_dg, s, cin)
    elif name == 'ngan_conv3x3_dgrad_pn':
        fn = lambda i: o.conv3x3_dgrad_pn(xout[i % n_buf], w_dg, s, leak, xin[i % n_buf], r)
    elif name == 'ngan_conv3x3_dbl':
        fn = lambda i: o.conv3x3_dbl(xin[i % n_buf], w_fwd, s, leak, xout[i % n_buf], r, xout[(i + 1) % n_buf])
    elif name == 'ngan_conv3x3_wgrad':
        dw = torch.zeros(cout, cin, 3, 3, device=dev)
        fn = lambda i: o.conv3x3_wgrad(xin[i % n_buf], xout[i % n_buf], s, dw, accumulate=False)
    else:
        return None
    for i in range(2):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float('inf')
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters * 1e3)
    return best


def family_roofline(shapes, hbm_peak, tf_peak, ms_step, label):
    """shapes: {(entry point, dims): launches per step}.  Every launch shape is graph-timed in isolation and weighted
    by its launches per step: family bytes / family time / peak."""
    tot_us = tot_bytes = tot_flops = 0.0
    n_launch = 0
    rows = []
    for (name, dims), cnt in sorted(shapes.items()):
        us = time_launch_isolated(name, dims)
        if us is None:
            continue
        by, fl = algorithmic_bytes(name, dims), conv_flops(name, dims)
        tot_us += cnt * us
        tot_bytes += cnt * by
        tot_flops += cnt * fl
        n_launch += cnt
        rows.append({'kernel': name, 'dims': list(dims), 'launches_per_step': cnt, 'us': round(us, 2),
                     'GBps': round(by / us / 1e3, 1), 'frac': round(by / us / 1e3 / hbm_peak, 4)})
    if not rows:
        return None, []
    ach = tot_bytes / tot_us / 1e3
    fam = {'family': label, 'launches_per_step': n_launch, 'launch_shapes': len(rows),
           'isolated_ms_per_step': round(tot_us / 1e3, 4),
           'share_of_step_time': round(tot_us / 1e3 / ms_step, 4),
           'algorithmic_bytes_per_step': int(tot_bytes), 'achieved': round(ach, 1), 'unit': 'GB/s',
           'frac': round(ach / hbm_peak, 4), 'tensor_TFLOPs': round(tot_flops / tot_us / 1e6, 1),
           'tensor_frac_of_sustained_peak': round(tot_flops / tot_us / 1e6 / tf_peak, 4),
           'timing': 'every launch shape graph-timed in isolation (20 launches over rotating buffers per graph, best '
                     'of 3 replays), weighted by its launches per step'}
    return fam, rows


def median_of_rounds(fn, rounds):
    vals = [fn() for _ in range(max(1, rounds))]
    return statistics.median(vals), [round(v, 4) for v in vals]


# ------------------------------------------------------------------------------------------------ CUDA arm: training step
def run_b200(args):
    import torch
    import torch.distributed as dist
    from neuron_gan_b200 import _lib
    from neuron_gan_b200.train_step import TrainStep, build_networks

    _lib.load()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import datetime
        # short collective timeout: a rank that falls out of step must fail in minutes, not hold the box
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=120))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak, hbm_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
    tf_peak = peaks.get('bf16_tflops_sustained', 1400.0)

    B, res, alpha = args.batch, args.res, args.alpha
    G, D = build_networks(res, alpha, seed=1, device=dev)
    step = TrainStep(G, D, use_graph=False if args.no_graph else None)
    # a small pool of different synthetic batches, pinned on the host and mirrored on the device; with data
    # parallelism every rank draws the GLOBAL batch's latents / epsilon on identically seeded CPU generators and keeps
    # its rows (dp.global_draws): the reproducible-stream path of SURVEY.md section 8d
    n_pool = 4
    host = [synthetic_images(B * world, res, seed=100 + i)[rank * B:(rank + 1) * B].contiguous().pin_memory()
            for i in range(n_pool)]
    devx = [h.to(dev) for h in host]
    torch.manual_seed(1234)
    draws = [tuple(t.to(dev) for t in step.draw_host(B)) for _ in range(n_pool)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled (nvidia-smi, every 20 ms) from before the warm-up to the end of the
    # end-to-end loop; the samples between the start of the timed region and the end of the e2e loop are reported
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # at least 3 untimed iterations: the second sight of a configuration captures its CUDA graph, replays start at 3
    for i in range(max(args.warmup, 3)):
        step(devx[i % n_pool], draws[i % n_pool])
    barrier()

    # ---- device-resident throughput
    launches0 = _lib.launch_count
    t0 = time.perf_counter()
    last = {}

    def resident_round():
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            last['stats'] = step(devx[i % n_pool], draws[i % n_pool])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ms, ms_rounds = median_of_rounds(resident_round, args.rounds)
    launches = (_lib.launch_count - launches0) // (args.steps * max(1, args.rounds))
    graph_replay = bool(step._graphs)
    stats = last['stats']
    last_stats = TrainStep.stats_dict(stats.cpu())
    TrainStep.check_nan(list(stats.cpu()))

    # ---- end to end through the public call: pinned host images -> H2D, CPU draws -> H2D, stats -> D2H
    # Host batches stream through a one-ahead prefetcher (H2D on a copy stream); the statistics of step i are read
    # back (pinned, async) and NaN-checked while step i+1 is already queued -- every copy is inside the timed region.
    from neuron_gan_b200.utils import DevicePrefetcher
    h2d = B * res * res * 4 + 3 * B * 512 * 4 + B * 4
    stat_bufs = [torch.empty(5, dtype=torch.float32).pin_memory() for _ in range(2)]
    n_e2e = [0]
    loader = DevicePrefetcher(lambda: (host[j % n_pool] for j in range(n_e2e[0])), dev,
                              gate=lambda: step.inputs_loaded)

    def run_e2e(n):
        n_e2e[0] = n
        pending = None
        for i, x in enumerate(loader):
            s = step(x)                      # draws z, z, eps, z on the CPU generator -> H2D
            hb = stat_bufs[i % 2]
            hb.copy_(s, non_blocking=True)   # the reference's six .item() calls, as one packed read
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                TrainStep.check_nan(pending[0].tolist())
            pending = (hb, ev)
        pending[1].synchronize()
        TrainStep.check_nan(pending[0].tolist())

    run_e2e(max(2, min(args.warmup, 4)))     # untimed: creates the staging buffers this path allocates lazily

    def e2e_round():
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        run_e2e(args.steps)
        f1.record()
        barrier()
        t = torch.tensor([f0.elapsed_time(f1) / args.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ms_e2e, e2e_rounds = median_of_rounds(e2e_round, args.rounds)
    clocks = sampler.stop(t0, time.perf_counter()) if rank == 0 else None
    peak_mem = torch.cuda.max_memory_allocated()

    # ---- roofline
    n_params = sum(p.numel() for net in (G, D) for p in net.active_parameters())
    step_bytes = B * MIN_BYTES_MB.get(res, 0.0) * 1e6
    adam_bytes = 28.0 * n_params
    nec_flops = B * NECESSARY_GFLOP.get(res, 0.0) * 1e9
    ach = step_bytes / (ms * 1e-3) / 1e9
    roofline = {'scope': 'step', 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak, 'peak_source': hbm_src,
                'unit': 'GB/s', 'frac': round(ach / hbm_peak, 4), 'traffic': None,
                'algorithmic_bytes_per_step': int(step_bytes),
                'algorithmic_bytes': f'{MIN_BYTES_MB.get(res)} MB per image (layer-granular minimum 15*B_D + 5*B_G, bf16 '
                                     f'features, SURVEY.md 8d) x {B} images per GPU; Adam (28 B x {n_params} active '
                                     f'parameters = {adam_bytes / 1e6:.0f} MB per step) is not included',
                'frac_with_adam_bytes': round((step_bytes + adam_bytes) / (ms * 1e-3) / 1e9 / hbm_peak, 4),
                'tensor_TFLOPs': round(nec_flops / (ms * 1e-3) / 1e12, 1),
                'tensor_frac_of_sustained_peak': round(nec_flops / (ms * 1e-3) / 1e12 / tf_peak, 4),
                'necessary_GFLOP_per_image': NECESSARY_GFLOP.get(res)}
    kernels = []
    if rank == 0 and not args.no_profile:
        # pass 1 (host-paced events around every launch, kernel by kernel on one stream): which launch shapes one
        # iteration consists of, and how often.  Its times are only used for the listing, not for the roofline.
        from neuron_gan_b200 import engine
        dp_was, step.dp = step.dp, False      # rank 0 alone runs this pass: no collective may be issued in it
        graph_was, step.use_graph = step.use_graph, False
        engine._Side.enabled = False
        fork_was, step.fork_chains = step.fork_chains, False
        _lib.start_profile()
        n_prof = 2
        for i in range(n_prof):
            step(devx[i % n_pool], draws[i % n_pool])
        prof = _lib.stop_profile()
        step.dp, step.use_graph, step.fork_chains = dp_was, graph_was, fork_was
        engine._Side.enabled = True
        agg = {}
        for name, a, dt in prof:
            k = (name, a[-5:] if name.startswith('ngan_conv3x3') else a[-4:])
            e = agg.setdefault(k, [0.0, 0])
            e[0] += dt
            e[1] += 1
        conv_shapes = {k: c // n_prof for k, (t, c) in agg.items() if k[0] in CONV_NAMES}
        wgrad_shapes = {k: c // n_prof for k, (t, c) in agg.items() if k[0] == 'ngan_conv3x3_wgrad'}
        # pass 2: every conv / wgrad launch shape graph-timed in isolation
        fam, rows = family_roofline(conv_shapes, hbm_peak, tf_peak, ms,
                                    'conv3x3 tcgen05 implicit GEMM (fwd / fwd_toim / dgrad / dgrad_pn / dbl)')
        wfam, wrows = family_roofline(wgrad_shapes, hbm_peak, tf_peak, ms, 'conv3x3 weight gradient (mma.sync + reduce)')
        roofline['family'], roofline['wgrad_family'] = fam, wfam
        allrows = sorted(rows + wrows, key=lambda r: -r['us'] * r['launches_per_step'])
        kernels = allrows[:8]
        big = [r for r in rows if r['us'] >= 20.0]
        if big:
            b = max(big, key=lambda r: r['frac'])
            traffic = None
            for f in ('r02_traffic.json', 'r01_traffic.json'):      # ncu --set full captures, per launch
                try:
                    tr = json.load(open(os.path.join(ROOT, 'profiles', f)))
                    traffic = traffic or tr.get(f'{b["kernel"]}|{"x".join(map(str, b["dims"]))}')
                except Exception:  # noqa: BLE001
                    pass
            fl = conv_flops(b['kernel'], b['dims'])
            roofline['best_launch'] = {'kernel': b['kernel'], 'dims': b['dims'], 'avg_launch_us': b['us'],
                                       'achieved': b['GBps'], 'frac': b['frac'], 'traffic': traffic,
                                       'algorithmic_bytes_per_launch': int(algorithmic_bytes(b['kernel'], b['dims'])),
                                       'tensor_TFLOPs': round(fl / b['us'] / 1e6, 1),
                                       'share_of_step_time': round(b['us'] * b['launches_per_step'] / 1e3 / ms, 4)}
        if args.dump_kernels:
            with open(args.dump_kernels, 'w') as fh:
                fh.write(f'# launch shapes of one iteration ({res}x{res}, {B} images), graph-timed in isolation\n')
                fh.write('# kernel,dims,launches_per_step,us,GBps,frac_of_hbm_peak\n')
                for r in allrows:
                    fh.write(f'{r["kernel"]},{"x".join(map(str, r["dims"]))},{r["launches_per_step"]},{r["us"]},'
                             f'{r["GBps"]},{r["frac"]}\n')
                fh.write('# other entry points (host-paced event pass: launches per step only)\n')
                for (name, dims), (t, c) in sorted(agg.items()):
                    if name not in CONV_NAMES and name != 'ngan_conv3x3_wgrad':
                        fh.write(f'{name},{"x".join(map(str, dims))},{c // n_prof}\n')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- baselines (rank 0, single-GPU runs only): the reference's own step, eager PyTorch on this GPU and on the host
    eager = None
    if world == 1 and not args.no_eager_baseline:
        try:
            torch.cuda.empty_cache()
            ips, ems, _, kind = time_reference_step(res, alpha, B, 5, 3, device=dev)
            eager = {'value': round(ips, 2), 'unit': 'images/s', 'ms_per_step': round(ems, 2), 'kind': kind,
                     'sample': f'5 timed + 3 warm-up iterations of the unmodified reference modules, eager PyTorch fp32 '
                               f'on this GPU (cudnn.benchmark, as train.py:128-144), batch {B}'}
        except Exception as e:  # noqa: BLE001
            eager = {'unavailable': str(e)[:200]}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or cpu_batch_for(res)
        n_it = 6 if res >= 256 else 10
        ips, cms, cores, kind = time_reference_step(res, alpha, cb, n_it, 2)
        what = 'unmodified reference modules' if kind == 'reference' else 'fp32 oracle port'
        cpu = {'value': round(ips, 3), 'unit': 'images/s', 'cores': cores, 'kind': kind,
               'sample': f'{n_it} timed + 2 warm-up iterations of the same step at batch {cb} ({what}, CPU)',
               'ms_per_step': round(cms, 1)}

    line = {'metric': METRIC, 'value': round(B * world / (ms * 1e-3), 2), 'unit': 'images/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms, 4), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'rounds': {'n': max(1, args.rounds), 'statistic': 'median', 'ms_per_step': ms_rounds,
                       'e2e_ms_per_step': e2e_rounds},
            'config': {'workload': workload_name(res, alpha), 'resolution': res, 'alpha': alpha, 'batch_per_gpu': B,
                       'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2': 'per-step activation working set (GBs) exceeds the 126 MB L2; 4 rotating input batches',
                       'launch': 'cuda-graph replay, wgrad kernels on forked streams' if graph_replay else 'eager',
                       'draws': 'global CPU draws sliced per rank (dp.global_draws order)' if world > 1 else 'CPU draws',
                       'losses': last_stats, 'peak_mem_GB': round(peak_mem / 2 ** 30, 2)},
            'e2e': {'value': round(B * world / (ms_e2e * 1e-3), 2), 'unit': 'images/s', 'ms_per_step': round(ms_e2e, 4),
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 20},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernels': kernels,
            'cpu_baseline': cpu, 'gpu_eager_baseline': eager}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CUDA arm: inference
def run_b200_eval(args):
    """BASELINE config 5: eval.py's path -- z (seeded CPU draw) -> G(z) under no_grad, chunked -> images on the HOST
    (the reference does .cpu() before writing the PNG grid, utils.py:583).  One 'step' = one gen_samples call."""
    import torch
    from neuron_gan_b200 import _lib
    from neuron_gan_b200.train_step import build_networks
    from neuron_gan_b200.utils import gen_samples_host
    _lib.load()
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak, hbm_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
    res = args.res
    G, _ = build_networks(res, 1.0, seed=1, device=dev)
    G.train(False)
    dtype = torch.bfloat16 if args.eval_dtype == 'bf16' else torch.float32
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    t_start = time.perf_counter()
    table = []
    launches0 = _lib.launch_count
    for n in [int(v) for v in args.eval_n.split(',')]:
        for _ in range(max(args.warmup, 3) if n <= 256 else 2):
            gen_samples_host(G, n, seed=0, dtype=dtype)
        torch.cuda.synchronize()
        reps = max(3, min(args.steps, 20)) if n <= 1024 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count
        e0.record()
        for _ in range(reps):
            imgs = gen_samples_host(G, n, seed=0, dtype=dtype)       # ends with the images in pinned host memory
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / reps
        launches = (_lib.launch_count - l0) // reps
        # device-resident: the same chunks without the device-to-host copy
        e0.record()
        for _ in range(reps):
            gen_samples_host(G, n, seed=0, dtype=dtype, to_host=False)
        e1.record()
        torch.cuda.synchronize()
        ms_dev = e0.elapsed_time(e1) / reps
        table.append({'n': n, 'images_per_s': round(n / ms_dev * 1e3, 1), 'e2e_images_per_s': round(n / ms_e2e * 1e3, 1),
                      'ms': round(ms_dev, 3), 'e2e_ms': round(ms_e2e, 3), 'launches': int(launches),
                      'd2h_bytes': int(imgs.numel() * imgs.element_size())})
    clocks = sampler.stop(t_start, time.perf_counter())
    head = max(table, key=lambda r: r['n'])
    by = EVAL_BYTES_MB_512 * 1e6 if res == 512 else None
    ach = head['images_per_s'] * by / 1e9 if by else None
    roofline = None
    if by:
        roofline = {'scope': 'generator forward pass', 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': hbm_peak,
                    'peak_source': hbm_src, 'unit': 'GB/s', 'frac': round(ach / hbm_peak, 4), 'traffic': None,
                    'algorithmic_bytes': f'{EVAL_BYTES_MB_512} MB per image (B_G: every layer pass reads its input and '
                                         'writes its output once, bf16; SURVEY.md 8d eval config)'}
    line = {'metric': EVAL_METRIC, 'value': head['images_per_s'], 'unit': 'images/s', 'n_gpus': 1,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': head['ms'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'pggan_{res}x{res}_generator_inference', 'n_images': head['n'],
                       'image_dtype': args.eval_dtype, 'chunk': 128,
                       'l2': 'per-chunk activations (GBs) exceed the 126 MB L2'},
            'e2e': {'value': head['e2e_images_per_s'], 'unit': 'images/s', 'ms_per_step': head['e2e_ms'],
                    'h2d_bytes_per_step': head['n'] * 512 * 4, 'd2h_bytes_per_step': head['d2h_bytes']},
            'gpu_launches': head['launches'], 'clocks': clocks, 'roofline': roofline, 'sweep': table,
            'cpu_baseline': None}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'eval':
        run_b200_eval(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
