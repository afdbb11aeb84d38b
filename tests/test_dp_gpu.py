"""Data-parallel equivalence on the CUDA path: 2 GPUs x B/2 == 1 GPU x B on the same global images and draws
(scripts/dp_equivalence.py under torchrun).  Needs two GPUs (`gpurun --gpus 2`); skipped on a one-GPU box -- the log
of the last run is kept in profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('res,alpha,batch', [(32, 0.5, 8), (128, 1.0, 8)])
def test_two_gpus_equal_one(res, alpha, batch):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'scripts', 'dp_equivalence.py'), str(res),
           str(alpha), str(batch), '4']
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert 'DP_EQUIVALENCE_OK' in out.stdout and 'replicas bit-identical after the DP run: True' in out.stdout
