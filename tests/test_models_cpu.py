"""Host-side logic of the reference-API mirror (no GPU): module tree / state-dict keys at every structural
state, same-seed initial weights, saved_attrs, checkpoint format round trip, error behaviour."""
import os

import pytest
import torch

from neuron_gan_b200 import models, utils
from oracle import pggan_oracle as O

ARCH = O.Arch()
GEN_F, DIS_F = [128, 64, 32, 32, 16, 16], [16, 16, 32, 32, 64, 128]


def build(res=16, alpha=1.0, seed=1):
    torch.manual_seed(seed)
    G = models.Generator_PG(list(GEN_F), image_size_init=16)
    D = models.Discriminator_PG(list(DIS_F), image_size_init=16)
    if res != 16:
        G.set_resolution(res, alpha)
        D.set_resolution(res, alpha)
    return G, D


def test_same_seed_gives_reference_weights(golden):
    G, D = build()
    gp, dp = O.build_params(ARCH, seed=1)
    gs, ds = G.state_dict(), D.state_dict()
    for name, key in O.g_key_map(1, False, ARCH).items():
        assert torch.equal(gs[key], gp[name]), name
        assert torch.equal(gs[key].flatten()[:8], golden['init']['g'][name]['head'])
    for name, key in O.d_key_map(1, False, ARCH).items():
        assert torch.equal(ds[key], dp[name]), name
        assert torch.equal(ds[key].flatten()[:8], golden['init']['d'][name]['head'])
    assert sum(p.numel() for p in G.parameters()) == 17093152
    assert sum(p.numel() for p in D.parameters()) == 494273


@pytest.mark.parametrize('key', ['r16_a1.0_b16', 'r32_a0.5_b4', 'r32_a1.0_b4', 'r64_a0.5_b64', 'r128_a1.0_b2',
                                 'r512_a0.5_b1', 'r512_a1.0_b2'])
def test_state_dict_keys_follow_the_reference(golden, key):
    ref = golden['cases'][key]
    G, D = build(ref['res'], ref['alpha'])
    assert sorted(G.state_dict().keys()) == ref['g_keys']
    assert sorted(D.state_dict().keys()) == ref['d_keys']
    assert G.image_size == D.image_size == ref['res']
    n = O.n_layers_for(ref['res'], ARCH)
    assert G.N_layers == D.N_layers == n
    # parameters the engine will differentiate == parameters the reference gives a gradient
    inv_g = {v: k for k, v in O.g_key_map(n, ref['alpha'] < 1, ARCH).items()}
    inv_d = {v: k for k, v in O.d_key_map(n, ref['alpha'] < 1, ARCH).items()}
    names_g = {id(p): inv_g[k] for k, p in G.named_parameters()}
    names_d = {id(p): inv_d[k] for k, p in D.named_parameters()}
    assert sorted(names_g[id(p)] for p in G.active_parameters()) == sorted(ref['g_grads'].keys())
    assert sorted(names_d[id(p)] for p in D.active_parameters()) == sorted(ref['d_grads'].keys())


def test_saved_attrs_and_progression():
    G, D = build()
    assert G.saved_attrs == ['LeakyReLU_neg_slope', 'N_colors', 'N_features_per_layer', 'N_layers', 'N_layers_max',
                             'image_size', 'image_size_init', 'image_size_max', 'latent_dim', 'training', 'alpha']
    assert D.saved_attrs == [a for a in G.saved_attrs if a != 'latent_dim']
    assert G.image_size_max == 512 and D.image_size_max == 512
    G.increase_resolution()
    assert float(G.alpha) == 0.0 and G.image_size == 32 and G.alpha_value() == 0.0
    with pytest.raises(AssertionError, match='The previous transition has not ended.'):
        G.increase_resolution()
    G.advance_transition(0.4)
    assert abs(G.alpha_value() - 0.4) < 1e-7 and len(G.conv_block_list) == 5
    G.advance_transition(0.6)
    assert G.alpha_value() >= 1.0 and len(G.conv_block_list) == 4 and len(G.layers) == 8
    with pytest.raises(AssertionError):
        G.set_resolution(48)
    D.set_resolution(64, 0.5)
    assert len(D.trunk_blocks()) == 1 and D.alpha_value() == 0.5 and len(D.FromIm_list) == 4
    # alpha set from outside (Checkpointer.set_saved_attrs) is seen by the host mirror
    D.alpha = torch.tensor(0.75)
    assert D.alpha_value() == 0.75


def test_no_cpu_fallback():
    G, D = build()
    with pytest.raises(RuntimeError, match='CUDA'):
        G(torch.randn(2, 512))
    with pytest.raises(RuntimeError, match='CUDA'):
        D(torch.randn(2, 1, 16, 16))
    with pytest.raises(RuntimeError, match='parameter holders'):
        G.layers[4](torch.randn(1, 128, 16, 16))


def test_checkpoint_roundtrip_reference_format(tmp_path):
    G, D = build(64, 0.5)
    path = os.path.join(tmp_path, 'GenDisc_test.pth')
    ck = utils.Checkpointer(G, D, 1e-4, path, N_epochs=10, verbose=False)
    ck.Loss_real[:3] = [1, 2, 3]
    ck.save_state(3)
    saved = torch.load(path, weights_only=False)
    assert sorted(saved.keys()) == sorted(['epoch', 'Generator_state', 'Generator_attrs', 'Discriminator_state',
                                           'Discriminator_attrs', 'lr', 'Loss_real', 'Loss_fake', 'Loss_G', 'Loss_D'])
    assert saved['Generator_attrs']['image_size'] == 64 and float(saved['Generator_attrs']['alpha']) == 0.5
    assert 'alpha' in saved['Discriminator_state'] and 'alpha' not in saved['Generator_state']
    G2, D2 = build(16, 1.0, seed=9)
    ck2 = utils.Checkpointer(G2, D2, 1e-4, path, N_epochs=10, verbose=False)
    ck2.load_state()
    assert ck2.epoch == 3 and list(ck2.Loss_real[:3]) == [1, 2, 3]
    assert G2.image_size == 64 and abs(G2.alpha_value() - 0.5) < 1e-7 and abs(D2.alpha_value() - 0.5) < 1e-7
    for (k, a), (_, b) in zip(G.state_dict().items(), G2.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(D.state_dict().items(), D2.state_dict().items()):
        assert torch.equal(a, b), k
    G3 = models.Generator_PG.from_state_dict(path, verbose=False)
    assert sorted(G3.state_dict().keys()) == sorted(G.state_dict().keys())


def test_latent_sampler_matches_oracle_and_memoises():
    torch.manual_seed(5)
    a = utils.sample_latent_vec((4, 512))
    torch.manual_seed(5)
    b = O.sample_latent((4, 512))
    assert torch.equal(a, b)
    state = torch.get_rng_state()
    s1 = utils.sample_latent_vec((3, 512), seed=0)
    assert torch.equal(torch.get_rng_state(), state)          # seeded draw leaves the global stream untouched
    assert utils.sample_latent_vec((3, 512), seed=0) is not None and torch.equal(s1, utils.sample_latent_vec((3, 512), seed=0))
    assert torch.allclose(a.norm(dim=1), torch.ones(4))
    with pytest.raises(ValueError):
        utils.sample_latent_vec((2, 4), mode='bogus')


def test_launcher_schedule_and_options_on_cpu():
    """Host logic of neuron_gan_b200/launch.py that needs no GPU: the learning-rate ramp (train.py:238-265; compared with
    the reference's own function in tests/test_oracle_vs_reference.py) and the option defaults (configs/config.py)."""
    import math
    from neuron_gan_b200.launch import LrSchedule, TrainConfig
    cfg = TrainConfig()
    assert (cfg.learning_rate, cfg.beta1, cfg.grad_pen_lambda, cfg.drift_epsilon) == (1e-4, 0.5, 10, 1e-3)
    assert cfg.transit_sch == [25000, 50000, 75000, 100000, 125000] and cfg.alpha_step == 1e-4 and cfg.batch_size == 8
    s = LrSchedule(1e-4, [10, 20], 30)
    assert s.value(0) == s.value(10) == s.value(20) == s.value(30) == 1e-4            # phase boundaries reset
    assert math.isclose(s.value(5), 1e-6, rel_tol=1e-9)                               # 1/100 at mid-phase
    assert math.isclose(s.value(1), 1e-4 * 0.01 ** (1 / 5), rel_tol=1e-12)
    assert s.value(6) is None and s.value(9) is None                                  # second half: left alone
    assert math.isclose(s.value(13), 1e-4 * 0.01 ** (3 / 5), rel_tol=1e-12)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-4)
    seen = []
    for epoch in range(0, 12):
        s.apply(opt, epoch)
        seen.append(opt.param_groups[0]['lr'])
    assert seen[5] == seen[6] == seen[9] and seen[10] == 1e-4 and seen[11] < 1e-4


def test_launcher_resume_restores_the_last_training_state(tmp_path):
    """launch.open_checkpoint (train.py:195-208) on CPU modules: a state saved mid fade-in at 32x32 comes back into
    freshly built 16x16 networks -- resolution, alpha, weights, loss series -- and training continues at epoch + 1."""
    from neuron_gan_b200 import models
    from neuron_gan_b200.launch import TrainConfig, open_checkpoint
    cfg = TrainConfig(N_gen_features=[32, 16, 16], N_dis_features=[16, 16, 32], image_size=64, N_epochs=20)

    def fresh(seed):
        torch.manual_seed(seed)
        return (models.Generator_PG(list(cfg.N_gen_features), image_size_init=16),
                models.Discriminator_PG(list(cfg.N_dis_features), image_size_init=16))

    G, D = fresh(1)
    for net in (G, D):
        net.increase_resolution()
        net.advance_transition(0.25)
    ckpt, first = open_checkpoint(cfg, G, D, str(tmp_path), resume=True, device=torch.device('cpu'))
    assert first == 1 and ckpt is not None                     # nothing to resume from yet
    ckpt.Loss_real[:6] = torch.arange(6).numpy()
    ckpt.save_state(6)
    G2, D2 = fresh(2)
    assert G2.image_size == 16 and not torch.equal(G2.state_dict()['layers.0.weight'], G.state_dict()['layers.0.weight'])
    ckpt2, first2 = open_checkpoint(cfg, G2, D2, str(tmp_path), resume=True, device=torch.device('cpu'))
    assert first2 == 7 and ckpt2.epoch == 6 and list(ckpt2.Loss_real[:6]) == [0, 1, 2, 3, 4, 5]
    assert G2.image_size == D2.image_size == 32 and abs(float(G2.alpha) - 0.25) < 1e-7 and abs(float(D2.alpha) - 0.25) < 1e-7
    for a, b in ((G, G2), (D, D2)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys())
        assert all(torch.equal(sa[k], sb[k]) for k in sa)
    # the epoch loop continues the statistic series of the checkpoint (adapt_critic looks at the epochs so far)
    from neuron_gan_b200.launch import restored_series
    names = ('D_loss', 'score_real', 'score_fake', 'G_loss', 'D_grad_pen')
    series = restored_series(ckpt2, first2, names)
    assert series['score_real'] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0] and len(series['score_fake']) == 6
    assert series['D_grad_pen'] == [] and restored_series(None, 1, names) == {k: [] for k in names}
    assert restored_series(ckpt2, 1, names) == {k: [] for k in names}
    _, first3 = open_checkpoint(cfg, *fresh(3), str(tmp_path), resume=False, device=torch.device('cpu'))
    assert first3 == 1
    assert open_checkpoint(cfg, G, D, None, True, torch.device('cpu')) == (None, 1)
