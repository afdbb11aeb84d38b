"""Live comparison oracle <-> UNMODIFIED reference; runs only where /root/reference exists (the
build container). On the GPU box these are skipped and the committed fixtures take over."""
import pytest
import torch

import ref_harness as rh
from oracle import pggan_oracle as O

pytestmark = pytest.mark.skipif(not rh.available(), reason='reference tree not present')
ARCH = O.Arch()


@pytest.mark.parametrize('res,alpha,batch', [(16, 1.0, 3), (32, 0.3, 2), (64, 1.0, 2), (128, 0.7, 1)])
def test_forward_and_losses(res, alpha, batch):
    G, D = rh.build_nets(res, alpha, seed=5)
    n = O.n_layers_for(res, ARCH)
    gp = O.g_state_to_named(G.state_dict(), n, alpha < 1, ARCH)
    dp = O.d_state_to_named(D.state_dict(), n, alpha < 1, ARCH)
    x = O.synthetic_images(batch, res, seed=11)
    z = O.sample_latent((batch, 512))
    with torch.no_grad():
        assert torch.allclose(G(z), O.g_forward(gp, z, n, alpha, ARCH), rtol=1e-5, atol=1e-6)
        assert torch.allclose(D(x), O.d_forward(dp, x, n, alpha, ARCH), rtol=1e-5, atol=1e-6)


def test_latent_sampler_bit_exact():
    _, _, ref_utils = rh.load()
    torch.manual_seed(123)
    a = ref_utils.sample_latent_vec((5, 512))
    torch.manual_seed(123)
    b = O.sample_latent((5, 512))
    assert torch.equal(a, b)


def test_small_architecture_iteration():
    """A non-default architecture (3 levels, 4x4 start) exercises the generic key maps."""
    gen_f, dis_f = [32, 16, 8], [8, 16, 32]
    arch = O.Arch(gen_features=gen_f, dis_features=dis_f, image_size=16)
    G, D = rh.build_nets(16, 0.5, gen_f, dis_f, image_size=16, seed=2)
    rng = torch.get_rng_state()
    x = O.synthetic_images(4, 16)
    stats, d_grads, g_grads, _ = rh.iteration(G, D, x)
    tr = O.Trainer(arch, seed=2, res=16, alpha=0.5)
    torch.set_rng_state(rng)
    mine = tr.iteration(x)
    for k in stats:
        assert abs(stats[k] - mine[k]) <= 2e-5 * max(1, abs(stats[k])), k
    dkm = O.d_key_map(3, True, arch)
    for name, key in dkm.items():
        if d_grads[key] is not None:
            assert torch.allclose(d_grads[key], tr.last_d_grads[name], rtol=1e-3, atol=1e-7), name


def test_pth_interchange_with_the_reference_classes(tmp_path):
    """Checkpoint layout (reference utils.py:142-223): a .pth written by this package's Checkpointer loads through
    the UNMODIFIED reference's from_state_dict / Checkpointer, and one written by the reference loads here -- same
    keys (also mid-transition, where the reference renumbers them), same tensors, same attrs."""
    import os
    from neuron_gan_b200 import models as my_models, utils as my_utils
    ref_models, _, ref_utils = rh.load()
    f = [128, 64, 32, 32, 16, 16]
    for res, alpha in ((64, 0.5), (128, 1.0)):
        # ours -> reference
        torch.manual_seed(3)
        G = my_models.Generator_PG(list(f), image_size_init=16)
        D = my_models.Discriminator_PG(list(reversed(f)), image_size_init=16)
        G.set_resolution(res, alpha)
        D.set_resolution(res, alpha)
        path = os.path.join(tmp_path, f'mine_{res}.pth')
        my_utils.Checkpointer(G, D, 1e-4, path, N_epochs=4, verbose=False).save_state(2)
        Gr = ref_models.Generator_PG.from_state_dict(path, verbose=False)
        Dr = ref_models.Discriminator_PG.from_state_dict(path, verbose=False)
        assert Gr.image_size == res and abs(float(Gr.alpha) - alpha) < 1e-7
        for net, ref in ((G, Gr), (D, Dr)):
            mine, theirs = net.state_dict(), ref.state_dict()
            assert list(mine.keys()) == list(theirs.keys())
            for k in mine:
                assert torch.equal(mine[k].cpu(), theirs[k].cpu()), k
        # reference -> ours
        Gr2, Dr2 = rh.build_nets(res, alpha, seed=4)
        path2 = os.path.join(tmp_path, f'ref_{res}.pth')
        ref_utils.Checkpointer(Gr2, Dr2, 1e-4, path2, N_epochs=4, verbose=False).save_state(1)
        G2 = my_models.Generator_PG.from_state_dict(path2, verbose=False)
        D2 = my_models.Discriminator_PG.from_state_dict(path2, verbose=False)
        for net, ref in ((G2, Gr2), (D2, Dr2)):
            mine, theirs = net.state_dict(), ref.state_dict()
            assert list(mine.keys()) == list(theirs.keys())
            for k in mine:
                assert torch.equal(mine[k].cpu(), theirs[k].cpu()), k
        assert G2.saved_attrs == Gr2.saved_attrs and D2.saved_attrs == Dr2.saved_attrs


def test_adaptive_critic_steps_match_the_reference():
    """utils.Calculate_D_steps (reference utils.py:105-120) on random score series, incl. the empty-series case."""
    import numpy as np
    from neuron_gan_b200 import utils as my_utils
    _, _, ref_utils = rh.load()
    assert my_utils.Calculate_D_steps([], [], 0, 5, 10) == ref_utils.Calculate_D_steps([], [], 0, 5, 10) == 5
    rng = np.random.RandomState(0)
    for trial in range(50):
        n = int(rng.randint(1, 40))
        real = list(rng.randn(n) * rng.rand() * 3)
        fake = list(rng.randn(n) * rng.rand() * 3 + rng.randn())
        n_min, n_max, period = int(rng.randint(0, 2)), int(rng.randint(2, 8)), int(rng.randint(1, 30))
        assert my_utils.Calculate_D_steps(real, fake, n_min, n_max, period) == \
            ref_utils.Calculate_D_steps(real, fake, n_min, n_max, period)


def test_lr_schedule_matches_the_reference_update_lr():
    """launch.LrSchedule against the UNMODIFIED `update_lr` of the reference, lifted out of train.py with ast
    (train.py is a script: importing it would start a training run) and executed with its own globals."""
    import ast
    import os
    import types
    import numpy as np
    import torch
    from neuron_gan_b200.launch import LrSchedule
    src = open(os.path.join(rh.REF_DIR, 'train.py')).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'update_lr')
    for transit_sch, n_epochs, lr0 in (([10, 20, 35], 50, 1e-4), ([7], 20, 3e-4), ([4, 9, 15, 22, 30], 41, 1e-3)):
        config = types.SimpleNamespace(learning_rate=lr0, transit_sch=transit_sch, N_epochs=n_epochs)
        boundaries = [0] + transit_sch + [n_epochs]                                  # train.py:238-243
        decay = [np.exp(np.log(1 / 100) / ((b - a) / 2)) for a, b in zip(boundaries[:-1], boundaries[1:])]
        env = {'config': config, 'transitions_epoch_boundaries': boundaries, 'lr_decay_rate': decay}
        exec(compile(ast.Module([fn], []), 'train.py', 'exec'), env)
        p = torch.nn.Parameter(torch.zeros(1))
        ref_opt, my_opt = torch.optim.Adam([p], lr=lr0), torch.optim.Adam([p], lr=lr0)
        sched = LrSchedule(lr0, transit_sch, n_epochs)
        for epoch in range(0, n_epochs + 1):
            env['update_lr'](ref_opt, epoch)
            sched.apply(my_opt, epoch)
            assert ref_opt.param_groups[0]['lr'] == my_opt.param_groups[0]['lr'], (transit_sch, epoch)


def test_generator_legacy_checkpoint_surgery_matches_the_reference(tmp_path):
    """from_state_dict on an OLD-format generator checkpoint (already-merged modules still in ToIm_list /
    conv_block_list, plus ToIm_prev.* and last_conv_block.* keys; reference models.py:411-436 with
    pop_state_dict_modules, models.py:37-63): same surviving keys and tensors as the reference's loader."""
    import os
    from collections import OrderedDict
    from neuron_gan_b200 import models as my_models
    ref_models, _, ref_utils = rh.load()
    torch.manual_seed(3)
    G = ref_models.Generator_PG([32, 16, 16, 8], image_size_init=16)
    G.set_resolution(64, 1.0)
    new_sd = G.state_dict()
    old_sd = OrderedDict()
    extra = 2                                     # two merged modules that the old format kept at the list heads
    for k, v in new_sd.items():
        for name in ('ToIm_list.', 'conv_block_list.'):
            if k.startswith(name):
                rest = k[len(name):]
                i, tail = rest.split('.', 1)
                k = f'{name}{int(i) + extra}.{tail}'
        old_sd[k] = v
    for i in range(extra):
        old_sd[f'ToIm_list.{i}.layers.0.weight'] = torch.randn(1, 8, 1, 1)
        old_sd[f'conv_block_list.{i}.1.weight'] = torch.randn(8, 8, 3, 3)
        old_sd[f'conv_block_list.{i}.4.weight'] = torch.randn(8, 8, 3, 3)
    old_sd['ToIm_prev.layers.0.weight'] = torch.randn(1, 16, 1, 1)
    old_sd['last_conv_block.1.weight'] = torch.randn(16, 16, 3, 3)
    old_sd['last_conv_block.4.weight'] = torch.randn(16, 16, 3, 3)
    path = os.path.join(tmp_path, 'old.pth')
    torch.save({'Generator_attrs': ref_utils.get_saved_attrs(G), 'Generator_state': old_sd}, path)
    theirs = ref_models.Generator_PG.from_state_dict(path, verbose=False).state_dict()
    mine = my_models.Generator_PG.from_state_dict(path, verbose=False).state_dict()
    assert list(mine.keys()) == list(theirs.keys()) == list(new_sd.keys())
    for k in theirs:
        assert torch.equal(mine[k], theirs[k]) and torch.equal(mine[k], new_sd[k]), k
