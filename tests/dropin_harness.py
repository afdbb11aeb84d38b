"""Builds a run directory in which the reference's UNMODIFIED train.py / eval.py execute (used by
tests/golden/gen_dropin_golden.py with the reference's own modules on CPU, and by tests/test_dropin_gpu.py with the
shim modules of this repository on the GPU).  Nothing of the reference is edited: files are copied as they are.

Run directory:  train.py eval.py configs/ data/  (reference)  +  models.py loss_functions.py utils.py  (reference's
own, or shim/)  +  configs/dropin_test.py (a user config, the mechanism configs/config.py:208-263 provides)  +  a
synthetic PNG dataset.  `sitecustomize.py` on PYTHONPATH seeds numpy's global generator, which the reference's
load-time noise fill (data/NeuronDataset.py:13-20) uses unseeded."""
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CONFIG = '''ID = 'dropin'
device = '{device}'
image_size = 64
N_gen_features = [128, 64, 32]
N_dis_features = [32, 64, 128]
transit_sch = [3, 7]
alpha_step = 0.5
N_epochs = 12
batch_size = 4
checkpointing_period = 4
learning_rate = 0.0001
image_preprocessing = 'device'
translation = 0.05
seed = 1
dataset_dir = r'{dataset_dir}'
'''


def make_dataset(directory, n=8, size=64):
    from PIL import Image
    os.makedirs(directory, exist_ok=True)
    rng = np.random.RandomState(3)
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n):
        a = 120 + 100 * np.sin(xx / (3.0 + i)) * np.cos(yy / (4.0 + i)) + 25 * rng.randn(size, size)
        Image.fromarray(np.clip(a, 1, 255).astype(np.uint8)).save(os.path.join(directory, f'im{i}.png'))


def build_run_dir(run, modules='shim', device='cuda'):
    """modules: 'shim' (this repository) or 'reference' (the reference's own models / loss_functions / utils)."""
    ref = ref_loader.ref_dir()
    if ref is None:
        raise RuntimeError('the reference scripts are not available (bash oracle/make_ref.sh in the build container)')
    os.makedirs(run, exist_ok=True)
    for f in ('train.py', 'eval.py'):
        shutil.copy(os.path.join(ref, f), run)
    for d in ('configs', 'data'):
        os.makedirs(os.path.join(run, d), exist_ok=True)
        for f in os.listdir(os.path.join(ref, d)):
            if f.endswith('.py'):
                shutil.copy(os.path.join(ref, d, f), os.path.join(run, d))
    src = os.path.join(ROOT, 'shim') if modules == 'shim' else ref
    for f in ('models.py', 'loss_functions.py', 'utils.py'):
        shutil.copy(os.path.join(src, f), run)
    dataset_dir = os.path.join(run, 'data', 'png')
    make_dataset(dataset_dir)
    with open(os.path.join(run, 'configs', 'dropin_test.py'), 'w') as fh:
        fh.write(CONFIG.format(device=device, dataset_dir=dataset_dir))
    site = os.path.join(run, '_site')
    os.makedirs(site, exist_ok=True)
    with open(os.path.join(site, 'sitecustomize.py'), 'w') as fh:
        fh.write('import numpy\nnumpy.random.seed(0)\n')
    return run


def run_script(run, args, timeout=900):
    """python <args> in the run directory; PYTHONPATH = sitecustomize dir, this repository, and (only for packages the
    image lacks) shim/stubs."""
    env = dict(os.environ)
    path = [os.path.join(run, '_site'), ROOT]
    for name in ('parse', 'matplotlib', 'skimage'):
        try:
            __import__(name)
        except ImportError:
            if ref_loader.STUBS_DIR not in path:
                path.append(ref_loader.STUBS_DIR)
    env['PYTHONPATH'] = os.pathsep.join(path)
    env['TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD'] = '1'        # SURVEY.md section 0 row 12 (the reference's bare torch.load)
    env['PYTHONDONTWRITEBYTECODE'] = '1'
    out = subprocess.run([sys.executable] + list(args), cwd=run, env=env, capture_output=True, text=True,
                         timeout=timeout, stdin=subprocess.DEVNULL)
    return out


def series_of(run):
    import torch
    ckpt = torch.load(os.path.join(run, 'weights', 'GenDisc_dropin.pth'), map_location='cpu', weights_only=False)
    return ckpt
