"""CPU: the oracle restatement reproduces the committed outputs of the reference itself (tests/golden/*.npz,
written by oracle/make_golden.py from /root/reference). Bit-exact: same torch CPU ops in the same order.
(torch CPU kernels are deterministic for a fixed thread count; allow 1e-6 slack for a different core count.)"""
import os

import numpy as np
import pytest
import torch

from oracle import CONFIGS
from oracle.make_golden import SAMPLER_CASES, SCHEDULE_CASES, DummyClassifier
from oracle.sampler import Sampler
from oracle.schedule import Tables
from oracle.unet import unet_forward, param_shapes
from oracle.weights import random_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('T,R,name', SCHEDULE_CASES)
def test_schedule_tables_bit_exact(T, R, name):
    g = np.load(os.path.join(GOLD, 'schedule.npz'))
    tab = Tables(T, R, name)
    key = 'T{}_R{}_{}'.format(T, R, name)
    assert np.array_equal(g[key + '/timestep_map'], tab.timestep_map)
    for n in Tables.NAMES:
        assert np.array_equal(g[key + '/' + n], getattr(tab, n)), n


def test_schedule_known_answers_from_survey():
    """SURVEY.md section 8(c) known-answer values (generated from the reference in the survey session)."""
    t = Tables(1000, 250, 'cosine')
    assert t.timestep_map[:3].tolist() == [2, 6, 10] and t.timestep_map[-2:].tolist() == [994, 998]
    assert int(t.timestep_map.sum()) == 125000
    assert t.betas[0] == 1.3841909524869855e-4 and t.betas[-1] == 0.9599992229055568
    assert t.alphas_cumprod[-1] == 2.4287669070348544e-6
    assert t.log_posterior_var_clipped[0] == t.log_posterior_var_clipped[1] == -9.322145735186036
    assert t.posterior_mean_coef_x0[1] == 0.6459776349131636
    assert t.sqrt_reciprocal_alphas_cumprod[-1] == 641.6631243437595
    t = Tables(1000, 25, 'cosine')
    assert int(t.timestep_map.sum()) == 12500 and t.betas[0] == 1.885771330707553e-3 and t.betas[-1] == 0.8960318437412871
    t = Tables(1000, 50, 'linear')
    assert int(t.timestep_map.sum()) == 25000 and t.betas[0] == 2.19342749215401e-3
    assert t.betas[-1] == 0.32735320663781853 and t.alphas_cumprod[-1] == 4.837047811818103e-5
    t = Tables(1000, 300, 'cosine')   # R does not divide T: 334 map entries, 333 betas (kept as the reference has it)
    assert len(t.timestep_map) == 334 and len(t.betas) == 333


@pytest.mark.parametrize('name', ['tiny', 'tiny_legacy', 'tiny_pool', 'tiny96', 'emnist', 'in64', 'in128'])
def test_unet_forward_matches_reference_output(name):
    cfg = CONFIGS[name]
    sd = random_state_dict(cfg, seed=1234)
    g = np.load(os.path.join(GOLD, 'unet_{}.npz'.format(name)))
    y = torch.from_numpy(g['y']) if 'y' in g.files else None
    out = unet_forward(cfg, sd, torch.from_numpy(g['x']), torch.from_numpy(g['t']), y)
    ref = torch.from_numpy(g['out'])
    assert float((out - ref).abs().max()) <= 1e-6 * float(ref.abs().max())


def test_timestep_embedding_known_answers():
    from oracle.unet import timestep_embedding
    e = timestep_embedding(torch.tensor([2]), 192)[0]
    assert np.allclose(e[:3].numpy(), [-0.416146844625473, -0.24375800788402557, -0.07992666214704514], atol=1e-7)
    assert np.allclose(e[96:99].numpy(), [0.9092974066734314, 0.9698361158370972, 0.9968007206916809], atol=1e-7)


def test_param_counts_match_survey():
    counts = {n: sum(int(np.prod(s)) for s in param_shapes(CONFIGS[n]).values()) for n in ('emnist', 'in64', 'in128')}
    assert counts == {'emnist': 17989442, 'in64': 295904454, 'in128': 421529606}
    assert len(param_shapes(CONFIGS['in64'])) == 541 and len(param_shapes(CONFIGS['emnist'])) == 309


@pytest.mark.parametrize('name', sorted(SAMPLER_CASES))
def test_sampler_trajectory_matches_reference(name):
    cname, dkw, use_clf = SAMPLER_CASES[name]
    cfg = CONFIGS[cname]
    sd = random_state_dict(cfg, seed=4321)
    g = np.load(os.path.join(GOLD, 'traj_{}.npz'.format(name)))
    clf = DummyClassifier(cfg['in_channels'], cfg['resolution'], cfg['num_classes'], seed=5) if use_clf else None
    s = Sampler(lambda x, t, y=None: unet_forward(cfg, sd, x, t, y), classifier=clf, **dkw)
    y = torch.from_numpy(g['y']) if 'y' in g.files else None
    noises = torch.from_numpy(g['noises'])
    traj = []
    out = s.denoise(torch.from_numpy(g['x_T']), {'y': y} if y is not None else {}, lambda t: noises[t], trajectory=traj)
    assert float((out - torch.from_numpy(g['final'])).abs().max()) <= 2e-5
    assert float((torch.stack([t[1] for t in traj]) - torch.from_numpy(g['traj_x'])).abs().max()) <= 2e-5
