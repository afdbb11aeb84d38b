"""The training launcher (neuron_gan_b200/launch.py): the reference's epoch control around TrainStep, fed by the device
image pipeline, through a resolution transition, with reference-format checkpoints."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_cfg(**kw):
    from neuron_gan_b200.launch import TrainConfig
    cfg = TrainConfig(N_gen_features=[128, 64], N_dis_features=[64, 128], image_size=32, batch_size=4, N_epochs=7,
                      transit_sch=[3], alpha_step=0.5, checkpointing_period=3)
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def test_epoch_control_through_a_transition(tmp_path):
    from neuron_gan_b200 import launch, models
    from neuron_gan_b200.data import NeuronImages
    from neuron_gan_b200.train_step import build_networks
    from neuron_gan_b200.utils import Checkpointer
    cfg = _small_cfg()
    G, D = build_networks(16, 1.0, seed=1, device='cuda', gen_features=cfg.N_gen_features,
                          dis_features=cfg.N_dis_features, image_size=cfg.image_size)
    images = NeuronImages(launch.synthetic_images(10, cfg.image_size), cfg.image_size, True, cfg.translation)
    path = str(tmp_path / 'GenDisc.pth')
    ckpt = Checkpointer(G, D, cfg.learning_rate, path, N_epochs=cfg.N_epochs, device='cuda', verbose=False)
    torch.manual_seed(3)
    lines = []
    hist = launch.pggan_train(cfg, images, G, D, ckpt, log=lines.append)
    assert [h['epoch'] for h in hist] == list(range(1, 8))
    # train.py:318-333: resolution doubles at epoch 3 with alpha 0, then +0.5 per epoch until the block is absorbed
    assert [h['image_size'] for h in hist] == [16, 16, 32, 32, 32, 32, 32]
    assert [round(h['alpha'], 3) for h in hist] == [1.0, 1.0, 0.0, 0.5, 1.0, 1.0, 1.0]
    # update_lr (train.py:250-265) is applied at the END of an epoch: epoch e trains with the rate set for e - 1
    sched = launch.LrSchedule(cfg.learning_rate, cfg.transit_sch, cfg.N_epochs)
    lr = cfg.learning_rate
    for h in hist:
        assert h['lr'] == pytest.approx(lr)
        new = sched.value(h['epoch'])
        lr = lr if new is None else new
    assert all(all(torch.isfinite(torch.tensor(h[k])) for k in ('D_loss', 'G_loss', 'score_real')) for h in hist)
    # checkpoints at epochs 3 and 6 in the reference's format, readable by the mirror of the reference's loader
    saved = torch.load(path, weights_only=False)
    assert saved['epoch'] == 6 and saved['Generator_attrs']['image_size'] == 32
    assert len(saved['Loss_real']) == 6 and saved['Loss_real'][5] == pytest.approx(hist[5]['score_real'])
    G2 = models.Generator_PG.from_state_dict(path, device=torch.device('cuda'), verbose=False)
    assert G2.image_size == 32
    assert torch.equal(G2.state_dict()['layers.0.weight'].cpu(), saved['Generator_state']['layers.0.weight'].cpu())


def test_adaptive_critic_in_the_loop():
    from neuron_gan_b200 import launch
    from neuron_gan_b200.data import NeuronImages
    from neuron_gan_b200.train_step import build_networks
    cfg = _small_cfg(n_critic=2, N_epochs=2, transit_sch=[], checkpointing_period=100)
    G, D = build_networks(16, 1.0, seed=1, device='cuda', gen_features=cfg.N_gen_features,
                          dis_features=cfg.N_dis_features, image_size=cfg.image_size)
    images = NeuronImages(launch.synthetic_images(8, cfg.image_size), cfg.image_size, True, cfg.translation)
    hist = launch.pggan_train(cfg, images, G, D, log=lambda s: None)
    assert [h['n_critic'] for h in hist] == [2, 2]


def test_command_line_entry_point():
    out = subprocess.run([sys.executable, '-m', 'neuron_gan_b200.launch', '--synthetic', '8', '--image_size', '32',
                          '--N_gen_features', '128', '64', '--N_dis_features', '64', '128', '--batch_size', '4',
                          '--N_epochs', '4', '--transit_sch', '2', '--alpha_step', '0.5'],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'done: epoch 4, 32x32, alpha 1.000' in out.stdout, out.stdout[-500:]


def test_similarity_loss_in_the_loop():
    """sim_loss_lambda > 0 (train.py:300, 343-348, 379-381): the term is added to the reported generator loss, decays
    with sim_loss_lambda_decay_rate, and -- having no gradient path into either network -- leaves the weights alone."""
    from neuron_gan_b200 import launch
    from neuron_gan_b200.data import NeuronImages
    from neuron_gan_b200.train_step import build_networks
    out = {}
    for key, lam, decay in (('off', 0.0, 0.0), ('on', 2.0, 0.0), ('decay', 2.0, 0.5)):
        cfg = _small_cfg(N_epochs=3, transit_sch=[], checkpointing_period=100, sim_loss_lambda=lam,
                         sim_loss_lambda_decay_rate=decay)
        G, D = build_networks(16, 1.0, seed=1, device='cuda', gen_features=cfg.N_gen_features,
                              dis_features=cfg.N_dis_features, image_size=cfg.image_size)
        images = NeuronImages(launch.synthetic_images(8, cfg.image_size), cfg.image_size, True, cfg.translation)
        torch.manual_seed(5)
        out[key] = (launch.pggan_train(cfg, images, G, D, log=lambda s: None), G.state_dict())
    h0, h1, h2 = out['off'][0], out['on'][0], out['decay'][0]
    assert all(h['G_sim_loss'] == 0 for h in h0) and all(h['G_sim_loss'] > 0 for h in h1)
    for e, (a, b, c) in enumerate(zip(h0, h1, h2)):
        assert b['G_loss'] == pytest.approx(a['G_loss'] + b['G_sim_loss'], abs=1e-6)
        # lambda = 2 * (1 - 0.5)^(epoch - 1): the same images and latents, so the term scales exactly with lambda
        assert c['G_sim_loss'] == pytest.approx(b['G_sim_loss'] * 0.5 ** e, rel=1e-5)
    for k, v in out['off'][1].items():      # no gradient path: the weights do not notice the term
        assert torch.equal(v, out['on'][1][k]), k
