"""Pin the oracle against the reference and (re)generate tests/golden/*.npz.

Runs ONLY in the build container (needs /root/reference). It imports the unmodified reference
(`nicediffusion.model`, `nicediffusion.diffusion`) with a stub for the missing `matplotlib`
(nicediffusion/utils.py:3 imports it at module top), then for every case:

  1. runs the reference on CPU fp32,
  2. runs the oracle restatement on the same weights / inputs / injected noise,
  3. asserts they agree BIT-EXACTLY (same torch ops in the same order), and
  4. stores the reference's outputs as a fixture.

    python -m oracle.make_golden          # from the repo root
"""
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def import_reference():
    for name in ('matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import nicediffusion.model as rm
    import nicediffusion.diffusion as rd
    return rm, rd


def make_inputs(cfg, batch, seed):
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal((batch, cfg['in_channels'], cfg['resolution'],
                                             cfg['resolution'])).astype(np.float32))
    t = torch.from_numpy(rs.randint(0, 1000, size=(batch,)).astype(np.int64))
    y = None
    if cfg.get('num_classes') is not None:
        y = torch.from_numpy(rs.randint(0, cfg['num_classes'], size=(batch,)).astype(np.int64))
    return x, t, y


class DummyClassifier(torch.nn.Module):
    """Deterministic stand-in for the noisy-image classifier (the reference ships none,
    utils.py:171-172): logits = W . [avgpool4(x) flattened, t/10]."""

    def __init__(self, in_ch, res, num_classes, seed):
        super().__init__()
        rs = np.random.RandomState(seed)
        feat = in_ch * (res // 4) ** 2 + 1
        self.w = torch.nn.Parameter(torch.from_numpy(
            (rs.standard_normal((num_classes, feat)) / np.sqrt(feat)).astype(np.float32)))

    def forward(self, x, t):
        f = torch.nn.functional.avg_pool2d(x, 4).flatten(1)
        f = torch.cat([f, t.float()[:, None] / 10.0], dim=1)
        return torch.tanh(f @ self.w.t()) * 3.0


SAMPLER_CASES = {
    # name: (cfg, diffusion kwargs, classifier?)
    'ddpm_interp': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                 sampling_var_type='learned_interpolation'), False),
    'ddpm_interp_cfg': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                     sampling_var_type='learned_interpolation',
                                     guidance_method='classifier_free', guidance_strength=0.8), False),
    'ddpm_learned': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='linear',
                                  sampling_var_type='learned'), False),
    'ddpm_small': ('tiny_legacy', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='linear',
                                       sampling_var_type='small'), False),
    'ddpm_large': ('tiny_legacy', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                       sampling_var_type='large'), False),
    'ddim_eta0': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='linear',
                               sampling_var_type='learned_interpolation', use_ddim=True, ddim_eta=0.0), False),
    'ddim_eta05_cfg': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                    sampling_var_type='learned_interpolation', use_ddim=True, ddim_eta=0.5,
                                    guidance_method='classifier_free', guidance_strength=1.5), False),
    'ddpm_classifier': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                     sampling_var_type='learned_interpolation',
                                     guidance_method='classifier', guidance_strength=2.0), True),
    'ddim_classifier': ('tiny', dict(original_num_steps=1000, rescaled_num_steps=10, beta_schedule='cosine',
                                     sampling_var_type='learned_interpolation', use_ddim=True, ddim_eta=0.0,
                                     guidance_method='classifier', guidance_strength=2.0), True),
}

SCHEDULE_CASES = [(1000, 250, 'cosine'), (1000, 25, 'cosine'), (1000, 50, 'linear'), (1000, 1000, 'linear'),
                  (1000, 300, 'cosine'), (100, 10, 'constant'), (4000, 100, 'cosine')]


def main(only_unet=None):
    """only_unet: list of config names -> regenerate just those UNet fixtures (skips schedules and trajectories)."""
    from oracle import CONFIGS
    from oracle.schedule import Tables
    from oracle.unet import unet_forward
    from oracle.weights import random_state_dict
    from oracle.sampler import Sampler

    rm, rd = import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    dev = torch.device('cpu')

    # ---- 1. schedules / respacing / tables ------------------------------------------------
    sched = {}
    unet_cases = (('tiny', 2), ('tiny_legacy', 2), ('tiny_pool', 2), ('tiny96', 2), ('emnist', 2), ('in64', 1), ('in128', 1))
    if only_unet:
        SCHEDULE_LOCAL, unet_cases = [], tuple(c for c in unet_cases if c[0] in only_unet)
    else:
        SCHEDULE_LOCAL = SCHEDULE_CASES
    dummy = rm.DiffusionModel(**{k: v for k, v in CONFIGS['tiny'].items()})
    for (T, R, name) in SCHEDULE_LOCAL:
        ref = rd.Diffusion(model=dummy, original_num_steps=T, rescaled_num_steps=R,
                           sampling_var_type='learned_interpolation', loss_type='hybrid',
                           beta_schedule=name, device=dev)
        tab = Tables(T, R, name)
        key = 'T{}_R{}_{}'.format(T, R, name)
        assert np.array_equal(ref.timestep_map.numpy(), tab.timestep_map), key
        sched[key + '/timestep_map'] = ref.timestep_map.numpy()
        for n in Tables.NAMES:
            a, b = getattr(ref, n), getattr(tab, n)
            assert a.dtype == np.float64 and np.array_equal(a, b), (key, n)
            sched[key + '/' + n] = a
        print('schedule', key, 'len', len(tab.betas), 'bit-exact')
    if not only_unet:
        np.savez_compressed(os.path.join(OUT, 'schedule.npz'), **sched)

    # ---- 2. UNet forward ---------------------------------------------------------------------
    for cname, batch in unet_cases:
        cfg = CONFIGS[cname]
        sd = random_state_dict(cfg, seed=1234)
        ref_model = rm.DiffusionModel(**cfg).eval()
        ref_model.load_state_dict(sd, strict=True)
        assert list(ref_model.state_dict().keys()) == list(sd.keys()), 'state-dict order differs: ' + cname
        x, t, y = make_inputs(cfg, batch, seed=99)
        with torch.no_grad():
            ref_out = ref_model(x, t, y) if y is not None else ref_model(x, t)
        ora_out = unet_forward(cfg, sd, x, t, y)
        assert torch.equal(ref_out, ora_out), cname
        print('unet', cname, 'B', batch, 'bit-exact; |out|max', float(ref_out.abs().max()))
        save = dict(x=x.numpy(), t=t.numpy(), out=ref_out.numpy())
        if y is not None:
            save['y'] = y.numpy()
        if cname == 'in64':  # keep the fixture small: one image, eps+var output only (98 KB)
            pass
        np.savez_compressed(os.path.join(OUT, 'unet_{}.npz'.format(cname)), **save)

    # ---- 3. sampler trajectories ------------------------------------------------------------
    for name, (cname, dkw, use_clf) in ({} if only_unet else SAMPLER_CASES).items():
        cfg = CONFIGS[cname]
        sd = random_state_dict(cfg, seed=4321)
        ref_model = rm.DiffusionModel(**cfg).eval()
        ref_model.load_state_dict(sd, strict=True)
        B = 2
        rs = np.random.RandomState(7)
        R = dkw['rescaled_num_steps']
        x_T = torch.from_numpy(rs.standard_normal((B, cfg['in_channels'], cfg['resolution'],
                                                   cfg['resolution'])).astype(np.float32))
        noises = torch.from_numpy(rs.standard_normal((R,) + tuple(x_T.shape)).astype(np.float32))
        y = None
        if cfg.get('num_classes') is not None:
            y = torch.from_numpy(rs.randint(1, cfg['num_classes'], size=(B,)).astype(np.int64))
        clf = DummyClassifier(cfg['in_channels'], cfg['resolution'], cfg['num_classes'], seed=5) if use_clf else None

        ref = rd.Diffusion(model=ref_model, loss_type='hybrid', classifier=clf, device=dev, **dkw)
        # inject noise: the reference draws torch.randn_like(x_t) once per step, t = R-1 ... 0
        order = list(reversed(range(R)))
        it = iter(order)
        real = torch.randn_like
        torch.randn_like = lambda ref_t, *a, **k: noises[next(it)].clone()
        try:
            ref_x = ref.denoise(x=x_T.clone(), kwargs={'y': y} if y is not None else None,
                                batch_size=B, progress=False)
        finally:
            torch.randn_like = real

        def model_fn(xx, tt, y=None, _cfg=cfg, _sd=sd):
            return unet_forward(_cfg, _sd, xx, tt, y)
        ora = Sampler(model_fn, classifier=clf, **dkw)
        traj = []
        ora_x = ora.denoise(x_T.clone(), {'y': y} if y is not None else {}, lambda t: noises[t], trajectory=traj)
        assert torch.equal(ref_x, ora_x), name
        print('sampler', name, 'bit-exact; final |x|max', float(ref_x.abs().max()))
        save = dict(x_T=x_T.numpy(), noises=noises.numpy(), final=ref_x.numpy(),
                    traj_x=np.stack([s[1].numpy() for s in traj]),
                    traj_x0=np.stack([s[2].numpy() for s in traj]))
        if y is not None:
            save['y'] = y.numpy()
        np.savez_compressed(os.path.join(OUT, 'traj_{}.npz'.format(name)), **save)

    # ---- 4. q-sample (diffuse) ------------------------------------------------------------------
    cfg = CONFIGS['tiny']
    ref_model = rm.DiffusionModel(**cfg).eval()
    ref = rd.Diffusion(model=ref_model, original_num_steps=1000, rescaled_num_steps=50,
                       sampling_var_type='learned_interpolation', loss_type='hybrid',
                       beta_schedule='cosine', device=dev)
    rs = np.random.RandomState(3)
    x0 = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 16, 16)).astype(np.float32))
    nz = torch.from_numpy(rs.standard_normal((2, 3, 16, 16)).astype(np.float32))
    ref_xt = ref.diffuse(x0, steps_to_do=20, noise=nz)
    ora = Sampler(None, 1000, 50, 'learned_interpolation', beta_schedule='cosine')
    assert torch.equal(ref_xt, ora.diffuse(x0, 20, nz))
    np.savez_compressed(os.path.join(OUT, 'diffuse.npz'), x0=x0.numpy(), noise=nz.numpy(), x_t=ref_xt.numpy())
    print('diffuse bit-exact')
    print('golden fixtures written to', OUT)


def main_big(only=None):
    """BASELINE-size trajectories (oracle/big_cases.py): the reference's step functions on the real ImageNet-64 /
    EMNIST / ImageNet-128 networks with injected noise; oracle asserted bit-exact at every step; the oracle re-run with
    fp16-rounded GEMM operands (free-running) is stored too -- it is the yardstick for the free-running bound of the
    16-bit operand mode, and running it on the GPU box's host would take minutes."""
    from oracle import CONFIGS
    from oracle.big_cases import BIG_CASES, WEIGHT_SEED, big_case_inputs
    from oracle.unet import unet_forward
    from oracle.weights import random_state_dict
    from oracle.sampler import Sampler

    rm, rd = import_reference()
    torch.set_num_threads(8)
    dev = torch.device('cpu')
    for name, (cname, dkw, B, start, n_steps) in BIG_CASES.items():
        if only and name not in only:
            continue
        cfg = CONFIGS[cname]
        sd = random_state_dict(cfg, seed=WEIGHT_SEED)
        ref_model = rm.DiffusionModel(**cfg).eval()
        ref_model.load_state_dict(sd, strict=True)
        inp = big_case_inputs(name)
        kwargs = {'y': inp['y']}
        ref = rd.Diffusion(model=ref_model, loss_type='hybrid', device=dev, **dkw)

        def model_fn(xx, tt, y=None, _cfg=cfg, _sd=sd):
            return unet_forward(_cfg, _sd, xx, tt, y)

        def model16(xx, tt, y=None, _cfg=cfg, _sd=sd):
            return unet_forward(_cfg, _sd, xx, tt, y, operand_cast=lambda v: v.half().float())
        ora, emu = Sampler(model_fn, **dkw), Sampler(model16, **dkw)
        if start == 'noise':
            x = inp['x_start'].clone()
        else:
            x = ref.diffuse(inp['x0_img'], steps_to_do=n_steps, noise=inp['q_noise'])
            assert torch.equal(x, ora.diffuse(inp['x0_img'], n_steps, inp['q_noise']))
        x_first = x.clone()
        ref_step = ref.ddim_denoising_step if dkw.get('use_ddim') else ref.denoising_step
        ora_step = ora.ddim_step if dkw.get('use_ddim') else ora.ddpm_step
        emu_step = emu.ddim_step if dkw.get('use_ddim') else emu.ddpm_step
        traj_x, traj_x0 = [], []
        xe = x.clone()
        real = torch.randn_like
        for i in range(n_steps):
            t = inp['t_first'] - i
            ts = (t * torch.ones(B)).to(dev)
            torch.randn_like = lambda ref_t, *a, _i=i, **k: inp['noises'][_i].clone()
            try:
                rx, rx0 = ref_step(x, t=ts, kwargs=kwargs)
            finally:
                torch.randn_like = real
            ox, ox0 = ora_step(x, ts, kwargs, inp['noises'][i])
            assert torch.equal(rx, ox) and torch.equal(rx0, ox0), (name, t)
            xe, _ = emu_step(xe, ts, kwargs, inp['noises'][i])
            traj_x.append(rx.clone()); traj_x0.append(rx0.clone())
            x = rx
            print('  ', name, 't', t, '|x|max', float(rx.abs().max()), 'emu16 err', float((xe - rx).abs().max()), flush=True)
        if start == 'diffused':
            # the public loop (diffusion.py:155-226) on the same inputs gives the same final image
            it = iter(range(n_steps))
            torch.randn_like = lambda ref_t, *a, **k: inp['noises'][next(it)].clone()
            try:
                loop = ref.denoise(x=x_first.clone(), kwargs=kwargs, batch_size=B, progress=False, steps_to_do=n_steps)
            finally:
                torch.randn_like = real
            assert torch.equal(loop, x), name
        save = dict(x_first=x_first.numpy(), final=x.numpy(), traj_x=np.stack([v.numpy() for v in traj_x]),
                    emu16_final=xe.numpy())
        if cname != 'in128':
            save['traj_x0'] = np.stack([v.numpy() for v in traj_x0])
        np.savez_compressed(os.path.join(OUT, 'bigtraj_{}.npz'.format(name)), **save)
        print('big trajectory', name, 'bit-exact; emu16 final err', float((xe - x).abs().max()), flush=True)


LOSS_CASES = {
    # name: (config, diffusion kwargs incl. loss_type)
    'hybrid_interp': ('tiny', dict(original_num_steps=100, rescaled_num_steps=100, beta_schedule='cosine',
                                   sampling_var_type='learned_interpolation', loss_type='hybrid')),
    'kl_learned': ('tiny', dict(original_num_steps=100, rescaled_num_steps=100, beta_schedule='linear',
                                sampling_var_type='learned', loss_type='KL')),
    'klr_interp': ('tiny', dict(original_num_steps=100, rescaled_num_steps=50, beta_schedule='cosine',
                                sampling_var_type='learned_interpolation', loss_type='KL_rescaled')),
    'simple_small': ('tiny_legacy', dict(original_num_steps=100, rescaled_num_steps=100, beta_schedule='linear',
                                         sampling_var_type='small', loss_type='simple')),
    'kl_large': ('tiny_legacy', dict(original_num_steps=100, rescaled_num_steps=100, beta_schedule='cosine',
                                     sampling_var_type='large', loss_type='KL')),
    'hybrid_emnist': ('emnist', dict(original_num_steps=1000, rescaled_num_steps=1000, beta_schedule='cosine',
                                     sampling_var_type='learned_interpolation', loss_type='hybrid')),
}


def loss_case_inputs(name):
    """Seeded inputs of a loss case (shared with the tests): x_0 in [-1, 1] with saturated pixels (the +-0.999 branches of
    the discretised likelihood), t including 0 (the NLL branch), labels, injected noise."""
    from oracle import CONFIGS
    from oracle.big_cases import hash_name
    cname, dkw = LOSS_CASES[name]
    cfg = CONFIGS[cname]
    rs = np.random.RandomState(hash_name('loss_' + name) % (2 ** 31))
    B = 4
    shape = (B, cfg['in_channels'], cfg['resolution'], cfg['resolution'])
    x0 = np.clip(rs.uniform(-1.3, 1.3, shape), -1, 1).astype(np.float32)
    noise = rs.standard_normal(shape).astype(np.float32)
    R = dkw['rescaled_num_steps']
    t = np.array([0, 1, R // 2, R - 1], dtype=np.int64)
    y = rs.randint(1, cfg['num_classes'], size=(B,)).astype(np.int64) if cfg.get('num_classes') else None
    return dict(x0=torch.from_numpy(x0), noise=torch.from_numpy(noise), t=torch.from_numpy(t),
                y=torch.from_numpy(y) if y is not None else None)


def main_loss():
    """Training loss + gradients of the reference (train mode, dropout 0 or seeded) on seeded inputs; the oracle
    (oracle/loss.py) asserted bit-exact for the loss values."""
    from oracle import CONFIGS
    from oracle.loss import LossOracle
    from oracle.unet import unet_forward
    from oracle.weights import random_state_dict
    rm, rd = import_reference()
    torch.set_num_threads(8)
    dev = torch.device('cpu')
    for name, (cname, dkw) in LOSS_CASES.items():
        cfg = dict(CONFIGS[cname])
        cfg['dropout'] = 0.0                      # dropout draws are not part of the parity contract
        sd = random_state_dict(cfg, seed=777)
        ref_model = rm.DiffusionModel(**cfg)
        ref_model.load_state_dict(sd, strict=True)
        ref = rd.Diffusion(model=ref_model, device=dev, **dkw)
        ref_model.train()
        inp = loss_case_inputs(name)
        kwargs = {'y': inp['y']} if inp['y'] is not None else None
        loss = ref.loss(x_0=inp['x0'], t=inp['t'], kwargs=kwargs, noise=inp['noise'])
        total = loss.mean()
        names = [n for n, _ in ref_model.named_parameters()]
        grads = torch.autograd.grad(total, list(ref_model.parameters()), allow_unused=True)
        # the model output the loss saw, and d loss.mean() / d out, for the kernel-level check
        x_t = ref.diffusion_step(x_0=inp['x0'], t=inp['t'], noise=inp['noise'])
        out = ref_model(x_t, ref.timestep_map[inp['t'].long()], **(kwargs or {})).detach().requires_grad_(True)
        okw = {k: v for k, v in dkw.items() if k != 'loss_type'}
        ora = LossOracle(lambda xx, tt, y=None: unet_forward(cfg, sd, xx, tt, y), dkw['loss_type'], **okw)
        l2 = ora.loss_from_output(inp['x0'], x_t, inp['t'], out, inp['noise'])
        assert torch.equal(l2.detach(), loss.detach()), (name, l2, loss)
        g_out = torch.autograd.grad(l2.mean(), out)[0]
        save = dict(loss=loss.detach().numpy(), out=out.detach().numpy(), x_t=x_t.numpy(), g_out=g_out.numpy())
        # a few parameter gradients (first / middle / last tensors) as end-to-end anchors
        pick = [0, len(names) // 2, len(names) - 2, len(names) - 1]
        for i in pick:
            if grads[i] is not None:
                save['grad/' + names[i]] = grads[i].numpy()
        np.savez_compressed(os.path.join(OUT, 'loss_{}.npz'.format(name)), **save)
        print('loss', name, 'bit-exact;', loss.detach().numpy())


if __name__ == '__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'loss':
        main_loss()
    elif len(sys.argv) > 1 and sys.argv[1] == 'big':
        main_big(only=sys.argv[2].split(',') if len(sys.argv) > 2 else None)
    else:
        main(only_unet=sys.argv[1].split(',') if len(sys.argv) > 1 else None)
