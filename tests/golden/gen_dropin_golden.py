"""Generate tests/golden/dropin_golden.pt: the UNMODIFIED reference train.py, run with the reference's OWN models /
loss_functions / utils on the CPU of the build container, 12 epochs through two resolution transitions
(16 -> 32 -> 64) on a synthetic PNG dataset (tests/dropin_harness.py).  Recorded: the four loss series of the final
checkpoint, its epoch / attrs / state-dict keys, and the monitoring line train.py printed at epoch 10.
tests/test_dropin_gpu.py runs the same script with the shim modules on the B200 and compares.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/gen_dropin_golden.py
"""
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import dropin_harness as H  # noqa: E402


def main():
    run = H.build_run_dir(tempfile.mkdtemp(prefix='dropin_ref_'), modules='reference', device='cpu')
    out = H.run_script(run, ['train.py', '--configs', 'dropin_test'])
    print(out.stdout[-3000:])
    if out.returncode != 0:
        print(out.stderr[-3000:])
        raise SystemExit(out.returncode)
    ckpt = H.series_of(run)
    mon = [l for l in out.stdout.splitlines() if l.startswith('Epoch:')]
    golden = {'epoch': ckpt['epoch'], 'Loss_real': ckpt['Loss_real'], 'Loss_fake': ckpt['Loss_fake'],
              'Loss_G': ckpt['Loss_G'], 'Loss_D': ckpt['Loss_D'], 'lr': ckpt['lr'],
              'Generator_attrs': {k: (float(v) if torch.is_tensor(v) else v) for k, v in ckpt['Generator_attrs'].items()},
              'Discriminator_attrs': {k: (float(v) if torch.is_tensor(v) else v)
                                      for k, v in ckpt['Discriminator_attrs'].items()},
              'g_keys': sorted(ckpt['Generator_state'].keys()), 'd_keys': sorted(ckpt['Discriminator_state'].keys()),
              'monitoring': mon, 'files': sorted(os.listdir(os.path.join(run, 'images', 'dropin')))}
    path = os.path.join(HERE, 'dropin_golden.pt')
    torch.save(golden, path)
    print('wrote', path, {k: golden[k] for k in ('epoch', 'Loss_real', 'Loss_D', 'monitoring', 'files')})


if __name__ == '__main__':
    main()
