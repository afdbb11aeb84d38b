#!/usr/bin/env python
"""Benchmark of the north-star path: full D+G WGAN-GP training iterations (reference train.py:350-394) of the
progressive-growing GAN, images/sec.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

One "step" = one iteration: critic step (D_W loss + gradient penalty, backward incl. double backward, Adam)
followed by the generator step (G_W loss, backward, Adam) on one batch of synthetic 1-channel images.
Default workload: the 512x512 final phase at 16 images per GPU (BASELINE.json config 3 is 16/GPU x 8 GPUs);
weak scaling keeps 16 images per GPU.  `--res/--alpha/--batch` select other phases for sweeps.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events over K steps with inputs resident in HBM;
`e2e` is the same K steps through the public TrainStep call with the images copied from pinned host memory,
the latent / epsilon draws made on the CPU generator and copied, and the statistics read back, every step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train images/sec (D+G WGAN-GP step)'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--res', type=int, default=512)
    ap.add_argument('--alpha', type=float, default=1.0)
    ap.add_argument('--batch', type=int, default=16, help='images per GPU')
    ap.add_argument('--cpu-batch', type=int, default=0, help='batch of the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='launch kernel by kernel instead of replaying a CUDA graph')
    ap.add_argument('--no-profile', action='store_true', help='skip the per-kernel timing pass (roofline)')
    ap.add_argument('--dump-kernels', default='', help='write the full per-kernel table of the timing pass here')
    return ap.parse_args()


def workload_name(res, alpha):
    phase = 'stable' if alpha >= 1 else f'fade-in(alpha={alpha:g})'
    return f'pggan_{res}x{res}_{phase}_wgan-gp_step'


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_batch_for(res):
    return {16: 16, 32: 16, 64: 16, 128: 8, 256: 4, 512: 2}.get(res, 2)


def time_cpu_port(res, alpha, batch, steps, warmup):
    """The oracle (CPU restatement of the reference algorithm, plain PyTorch fp32) on all host cores."""
    import torch
    from oracle import pggan_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    tr = O.Trainer(O.Arch(), seed=1, res=res, alpha=alpha)
    x = O.synthetic_images(batch, res)
    for _ in range(warmup):
        tr.iteration(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        tr.iteration(x)
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = args.cpu_batch or cpu_batch_for(args.res)
    # bound the sample so the whole run stays within minutes on the box's host cores
    steps, warmup = min(args.steps, 5), min(args.warmup, 2)
    ips, ms, cores = time_cpu_port(args.res, args.alpha, batch, steps, max(warmup, 1))
    sample = f'{steps} timed + {max(warmup, 1)} warm-up iterations of the same step at batch {batch} (fp32, CPU)'
    line = {'impl': 'reference', 'metric': METRIC, 'value': ips, 'unit': 'images/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': max(warmup, 1), 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload_name(args.res, args.alpha), 'resolution': args.res, 'alpha': args.alpha,
                       'batch': batch, 'device': 'cpu'},
            'cpu_baseline': {'value': ips, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
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


def algorithmic_bytes(name, a):
    """Minimum HBM bytes one launch must move (inputs + outputs once, bf16 features, fp32 scales/images)."""
    if name == 'ngan_conv3x3_fwd':
        B, ci, co, H, W = a[-5:]
        return B * H * W * (2 * ci + 2 * co + 4)
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


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from neuron_gan_b200 import _lib
    from neuron_gan_b200.train_step import TrainStep, build_networks
    from oracle import pggan_oracle as O      # only synthetic_images + the cpu_baseline leg below

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
    # a small pool of different synthetic batches, pinned on the host and mirrored on the device
    n_pool = 4
    host = [O.synthetic_images(B, res, seed=100 + rank * 17 + i).pin_memory() for i in range(n_pool)]
    devx = [h.to(dev) for h in host]
    draws = [step.draw(B, dev) for _ in range(n_pool)]
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
    mem0 = torch.cuda.max_memory_allocated()
    launches0 = _lib.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        stats = step(devx[i % n_pool], draws[i % n_pool])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (_lib.launch_count - launches0) // args.steps
    graph_replay = bool(step._graphs)
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
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    run_e2e(args.steps)
    f1.record()
    barrier()
    clocks = sampler.stop(t0, time.perf_counter()) if rank == 0 else None
    ms_e2e = f0.elapsed_time(f1) / args.steps

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- per-kernel timing pass (events around every launch; perturbs throughput, so it is separate)
    roofline, kernels = None, []
    if rank == 0 and not args.no_profile:
        from neuron_gan_b200 import engine
        dp_was, step.dp = step.dp, False      # rank 0 alone runs this pass: no collective may be issued in it
        graph_was, step.use_graph = step.use_graph, False     # kernel by kernel, on one stream, events around each
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
            e = agg.setdefault(k, [0.0, 0, name, a])
            e[0] += dt
            e[1] += 1
        total = sum(e[0] for e in agg.values())
        top = sorted(agg.values(), key=lambda e: -e[0])
        if args.dump_kernels:
            fam = {}
            for tot, cnt, name, a in top:
                f = fam.setdefault(name, [0.0, 0])
                f[0] += tot
                f[1] += cnt
            with open(args.dump_kernels, 'w') as fh:
                fh.write(f'# per-kernel CUDA-event timing pass, {n_prof} steps, total {total / n_prof:.3f} ms/step\n')
                fh.write('# family,share,ms_per_step,launches_per_step\n')
                for name, (tot, cnt) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
                    fh.write(f'{name},{tot / total:.4f},{tot / n_prof:.4f},{cnt // n_prof}\n')
                fh.write('# kernel,dims,share,launches_per_step,avg_us,GBps,TFLOPs\n')
                for tot, cnt, name, a in top:
                    by, fl = algorithmic_bytes(name, a), conv_flops(name, a)
                    avg_s = tot / cnt * 1e-3
                    fh.write(f'{name},{"x".join(map(str, a[-5:]))},{tot / total:.4f},{cnt // n_prof},{avg_s * 1e6:.1f},'
                             f'{by / avg_s / 1e9 if by else 0:.0f},{fl / avg_s / 1e12 if fl else 0:.1f}\n')
        for tot, cnt, name, a in top[:8]:
            by, fl = algorithmic_bytes(name, a), conv_flops(name, a)
            avg_s = tot / cnt * 1e-3
            kernels.append({'kernel': name, 'dims': list(a[-5:]), 'share': round(tot / total, 4),
                            'launches_per_step': cnt // n_prof, 'avg_us': round(avg_s * 1e6, 2),
                            'GBps': round(by / avg_s / 1e9, 1) if by else None,
                            'TFLOPs': round(fl / avg_s / 1e12, 2) if fl else None})
        # dominant kernel = the CUDA kernel with the largest share of the step (the tcgen05 implicit-GEMM conv serves
        # the forward, data-gradient, PixelNorm-backward and double-backward entry points); the roofline is reported
        # for its heaviest launch shape
        conv_names = ('ngan_conv3x3_fwd', 'ngan_conv3x3_fwd_toim', 'ngan_conv3x3_dgrad', 'ngan_conv3x3_dgrad_pn',
                      'ngan_conv3x3_dbl')
        fam_of = lambda n: 'conv3x3 tcgen05 implicit GEMM (fwd/dgrad/dgrad_pn/dbl)' if n in conv_names else n
        fam_tot = {}
        for tot, cnt, name, a in top:
            fam_tot[fam_of(name)] = fam_tot.get(fam_of(name), 0.0) + tot
        dom = max(fam_tot, key=fam_tot.get)
        # ... = the launch shape with the longest single launch: the event pair around a launch also times ~5-10 us of
        # launch latency on an idle stream (this pass is host-paced), which distorts the short launches
        cand = sorted((e for e in top if fam_of(e[2]) == dom and algorithmic_bytes(e[2], e[3])),
                      key=lambda e: -e[0] / e[1])
        if cand:
            tot, cnt, name, a = cand[0]
            by, fl = algorithmic_bytes(name, a), conv_flops(name, a)
            avg_s = tot / cnt * 1e-3
            ach = by / avg_s / 1e9
            traffic = None
            try:
                tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))
                traffic = tr.get(f'{name}|{"x".join(map(str, a[-5:]))}')
            except Exception:  # noqa: BLE001
                pass
            roofline = {'kernel': name, 'dims': list(a[-5:]), 'family': dom,
                        'family_share_of_step': round(fam_tot[dom] / total, 4), 'bound': 'hbm',
                        'achieved': round(ach, 1), 'peak': hbm_peak, 'peak_source': hbm_src, 'unit': 'GB/s',
                        'frac': round(ach / hbm_peak, 4), 'traffic': traffic, 'share_of_step': round(tot / total, 4),
                        'tensor_TFLOPs': round(fl / avg_s / 1e12, 2) if fl else None,
                        'tensor_frac_of_sustained_peak': round(fl / avg_s / 1e12 / tf_peak, 4) if fl else None,
                        'algorithmic_bytes_per_launch': by, 'avg_launch_us': round(avg_s * 1e6, 2)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cb = args.cpu_batch or cpu_batch_for(res)
        n_it = 6 if res >= 256 else 10
        ips, cms, cores = time_cpu_port(res, alpha, cb, n_it, 2)
        cpu = {'value': round(ips, 3), 'unit': 'images/s', 'cores': cores, 'kind': 'port',
               'sample': f'{n_it} timed + 2 warm-up iterations of the same step at batch {cb} (fp32 oracle, CPU)',
               'ms_per_step': round(cms, 1)}

    line = {'metric': METRIC, 'value': round(B * world / (ms * 1e-3), 2), 'unit': 'images/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms, 4), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': workload_name(res, alpha), 'resolution': res, 'alpha': alpha, 'batch_per_gpu': B,
                       'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2': 'per-step activation working set (GBs) exceeds the 126 MB L2; 4 rotating input batches',
                       'launch': 'cuda-graph replay, wgrad kernels on a forked stream' if graph_replay else 'eager',
                       'losses': last_stats, 'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)},
            'e2e': {'value': round(B * world / (ms_e2e * 1e-3), 2), 'unit': 'images/s', 'ms_per_step': round(ms_e2e, 4),
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 20},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernels': kernels,
            'cpu_baseline': cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
