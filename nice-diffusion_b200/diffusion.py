"""`Diffusion` -- drop-in for nicediffusion.diffusion.Diffusion (reference diffusion.py:22-369) on ndb200 kernels.

Same constructor, public numpy tables and `timestep_map`, `denoise / diffuse / diffusion_step /
get_eps_and_log_var / denoising_step / ddim_denoising_step` with the reference's signatures, return values and
assertion behaviour. Underneath:

  * the schedule / respacing / posterior tables are computed on the host in numpy fp64 with the reference's
    arithmetic (bit-exact, tests/test_schedule.py) and uploaded ONCE as fp32 rows (the reference re-uploads a
    table on every `extract`, diffusion.py:492);
  * each step is one fused kernel (`ndb_sampler_step`) after the UNet forward(s);
  * `denoise` replays ONE CUDA graph per step (UNet plan + step kernel + device-side step counter), with the
    classifier-free-guidance pair evaluated as a doubled batch [y ; 0] in a single forward;
  * noise: `noise='torch'` (default) draws `torch.randn_like` per step like the reference; `noise='philox'`
    generates it inside the step kernel keyed by (seed, global sample id, step) so that sharding the batch over
    GPUs does not change any sample; an explicit `[R,B,C,H,W]` tensor injects noise (parity tests).

Training: `loss` (fused loss + gradient kernel over a differentiable model forward) and `trainer.Trainer`
(data-parallel gradient exchange over NCCL), SURVEY.md section 8 f-1 / e-2.
"""
import math

import numpy as np
import torch
import tqdm

from . import _lib as L
from . import ops
from .engine import capture_guard


def get_beta_schedule(schedule_method, num_steps, beta_0, beta_T):
    """diffusion.py:445-475 (numpy fp64; bit-exact)."""
    if schedule_method == 'linear':
        return np.linspace(beta_0, beta_T, num_steps, dtype=np.float64)
    elif schedule_method == 'constant':
        return beta_0 * np.ones(num_steps, dtype=np.float64)
    elif schedule_method == 'cosine':
        def f(t):
            return math.cos((t + 0.008) / (1.0 + 0.008) * math.pi / 2) ** 2
        return np.array([min(1 - f((s + 1) / num_steps) / f(s / num_steps), 0.999) for s in range(num_steps)])
    raise NotImplementedError("unimplemented variance scheduling method: {}".format(schedule_method))


_VAR_TYPES = ('small', 'large', 'learned', 'learned_interpolation')
_LOSS_TYPES = ('simple', 'KL', 'KL_rescaled', 'hybrid')


class Diffusion:
    def __init__(self, model, original_num_steps, rescaled_num_steps, sampling_var_type, loss_type,
                 betas=None, beta_schedule='linear', guidance_method=None, guidance_strength=None, classifier=None,
                 use_ddim=False, ddim_eta=None, device=None):
        self.model = model
        if device is None:
            self.device = torch.device('cuda') if torch.cuda.is_available() else torch.device('cpu')
        else:
            self.device = device
        self.model.to(self.device)
        self.model.eval()

        self.guidance = guidance_method
        if guidance_method not in (None, 'classifier', 'classifier_free'):
            raise NotImplementedError(guidance_method)
        assert guidance_method is None or self.model.conditional, 'can only use guidance if model is conditional'
        self.strength = guidance_strength
        self.classifier = classifier
        if self.classifier is not None:
            self.classifier.float().to(self.device)
            self.classifier.eval()

        self.original_num_steps = original_num_steps
        self.rescaled_num_steps = rescaled_num_steps
        if sampling_var_type not in _VAR_TYPES:
            raise NotImplementedError(sampling_var_type)
        if loss_type not in _LOSS_TYPES:
            raise NotImplementedError(loss_type)
        self.sampling_var_type = sampling_var_type
        self.loss_type = loss_type
        if use_ddim:
            assert ddim_eta is not None, 'please supply eta if you want to use ddim'
        self.use_ddim = use_ddim
        self.ddim_eta = ddim_eta

        T, R = original_num_steps, rescaled_num_steps
        if betas is None:
            betas = get_beta_schedule(beta_schedule, T, 0.0001 * 1000 / T, 0.02 * 1000 / T)
        else:
            assert len(betas) == T, 'betas must be the right length!'
            betas = np.array(betas, dtype=np.float64)
        # respacing (diffusion.py:94-107): keep every (T//R)-th step starting at T//(2R); new betas from the
        # ratios of the kept cumulative products. If R does not divide T the list is not of length R (kept).
        abar_full = np.cumprod(1.0 - betas, axis=0)
        kept = list(range(T // (2 * R), T + T // (2 * R), T // R))
        kept_lookup = set(kept)
        prev, new_betas = 1.0, []
        for i, a in enumerate(abar_full):
            if i in kept_lookup:
                new_betas.append(1.0 - a / prev)
                prev = a
        betas = np.array(new_betas)
        assert (betas > 0).all() and (betas <= 1).all(), 'betas in invalid range'

        self.betas = betas
        self.timestep_map = torch.tensor(kept, device=self.device, dtype=torch.long)
        alphas = 1.0 - betas
        self.alphas_cumprod = abar = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = abar_prev = np.append(1.0, abar[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(abar)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - abar)
        self.sqrt_reciprocal_alphas_cumprod = np.sqrt(1.0 / abar)
        self.sqrt_reciprocal_alphas_minus_one_cumprod = np.sqrt(1.0 / abar - 1)
        self.posterior_mean_coef_x0 = np.sqrt(abar_prev) * betas / (1.0 - abar)
        self.posterior_mean_coef_xt = np.sqrt(alphas) * (1.0 - abar_prev) / (1.0 - abar)
        self.posterior_variance = betas * (1.0 - abar_prev) / (1.0 - abar)
        self.log_posterior_var_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))

        # sampling options that are additions to the reference API
        self.noise = 'torch'      # 'torch' | 'philox'
        self.seed = 0             # philox seed
        self.sample_ids = None    # LongTensor [B] of global sample indices (philox keying); default arange(B)
        # classifier guidance: capture the classifier's forward + input-gradient inside the per-step CUDA graph (between the
        # UNet plan and the fused step kernel). False: reference-shaped host loop over the public step functions.
        self.classifier_in_graph = True
        self._tables_dev = {}     # sampling_var_type -> device tables (the fixed-variance row depends on it)

    # ------------------------------------------------------------------------------------------
    # device tables
    # ------------------------------------------------------------------------------------------
    def _var_code(self):
        return {'small': L.VAR_FIXED, 'large': L.VAR_FIXED, 'learned': L.VAR_LEARNED,
                'learned_interpolation': L.VAR_LEARNED_INTERP}[self.sampling_var_type]

    def _tables(self):
        """[TAB_COUNT][n] fp32 on the device: each fp64 table cast to fp32 exactly as `extract` does."""
        if self.sampling_var_type not in _VAR_TYPES:
            raise NotImplementedError(self.sampling_var_type)
        cached = self._tables_dev.get(self.sampling_var_type)
        if cached is None:
            n = len(self.betas)
            if self.sampling_var_type == 'large':
                logvar_min = np.log(np.append(self.posterior_variance[1], self.betas[1:]))   # diffusion.py:259
            elif self.sampling_var_type == 'small':
                logvar_min = np.log(np.maximum(self.posterior_variance, 1e-20))               # diffusion.py:261
            else:
                logvar_min = self.log_posterior_var_clipped
            rows = [None] * L.TAB_COUNT
            rows[L.TAB_SQRT_RECIP_ABAR] = self.sqrt_reciprocal_alphas_cumprod
            rows[L.TAB_SQRT_RECIPM1_ABAR] = self.sqrt_reciprocal_alphas_minus_one_cumprod
            rows[L.TAB_POST_COEF_X0] = self.posterior_mean_coef_x0
            rows[L.TAB_POST_COEF_XT] = self.posterior_mean_coef_xt
            rows[L.TAB_LOGVAR_MIN] = logvar_min
            rows[L.TAB_LOGVAR_MAX] = np.log(self.betas)
            rows[L.TAB_ABAR] = self.alphas_cumprod
            rows[L.TAB_ABAR_PREV] = self.alphas_cumprod_prev
            rows[L.TAB_SQRT_1M_ABAR] = self.sqrt_one_minus_alphas_cumprod
            rows[L.TAB_LOG_POST_VAR_CLIPPED] = self.log_posterior_var_clipped
            tab = np.stack([np.asarray(r, dtype=np.float64) for r in rows]).reshape(L.TAB_COUNT, n)
            cached = self._tables_dev[self.sampling_var_type] = dict(
                tab=torch.from_numpy(tab).float().to(self.device).contiguous(),
                sqrt_abar=torch.from_numpy(self.sqrt_alphas_cumprod).float().to(self.device),
                sqrt_1m_abar=torch.from_numpy(self.sqrt_one_minus_alphas_cumprod).float().to(self.device),
                tmap_f32=self.timestep_map.float().contiguous())
        return cached

    def _check_step_index(self, t, what):
        """The kernels index the tables with t unchecked; the reference's `extract` (torch.gather) raises for an index
        outside [0, len(betas)) -- so does this, on the host."""
        n = len(self.betas)
        lo, hi = (int(t), int(t)) if not torch.is_tensor(t) else (int(t.min().item()), int(t.max().item()))
        if lo < 0 or hi >= n:
            raise IndexError('{}: step index {} is out of bounds for a schedule of {} respaced steps'.format(
                what, lo if lo < 0 else hi, n))

    # ------------------------------------------------------------------------------------------
    # forward process
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def diffuse(self, x_0, steps_to_do=None, noise=None):
        """diffusion.py:132-153."""
        if steps_to_do is None or steps_to_do > self.rescaled_num_steps:
            steps_to_do = self.rescaled_num_steps
        self._check_step_index(steps_to_do - 1, 'diffuse')
        timestep = ((steps_to_do - 1) * torch.ones(x_0.shape[0])).to(self.device)
        return self.diffusion_step(x_0, t=timestep, noise=noise)

    def diffusion_step(self, x_0, t, noise=None):
        """q(x_t | x_0), diffusion.py:232-240."""
        if noise is None:
            noise = torch.randn_like(x_0).to(self.device)
        tabs = self._tables()
        self._check_step_index(t, 'diffusion_step')
        x_0 = x_0.to(self.device, torch.float32).contiguous()
        noise = noise.to(self.device, torch.float32).contiguous()
        return ops.qsample(x_0, noise, t.to(self.device).float().contiguous(), tabs['sqrt_abar'], tabs['sqrt_1m_abar'])

    # ------------------------------------------------------------------------------------------
    # reverse process: per-step API (same signatures / returns as the reference)
    # ------------------------------------------------------------------------------------------
    def get_eps_and_log_var(self, x_t, t, kwargs):
        """diffusion.py:242-264 -- returned as separate tensors for API parity (the sampling loop itself never
        materialises log_var; it is folded into the step kernel)."""
        out = self.model(x_t, self.timestep_map[t.long()], **(kwargs or {}))
        tabs = self._tables()['tab']
        ti = t.long()
        shape = (x_t.shape[0],) + (1,) * (x_t.dim() - 1)
        if self.sampling_var_type in ('learned', 'learned_interpolation'):
            eps, v = torch.split(out, int(out.shape[1] / 2), dim=1)
            if self.sampling_var_type == 'learned':
                return eps, v
            min_log = tabs[L.TAB_LOGVAR_MIN][ti].reshape(shape)
            max_log = tabs[L.TAB_LOGVAR_MAX][ti].reshape(shape)
            frac = (v + 1) / 2
            return eps, frac * max_log + (1 - frac) * min_log
        return out, tabs[L.TAB_LOGVAR_MIN][ti].reshape(shape).expand(x_t.shape)

    def _model_pair(self, x_t, t, kwargs):
        """UNet output for the step at respaced index tensor `t`; with classifier-free guidance the conditional
        and the null-class (label 0, diffusion.py:281) evaluations run as ONE doubled batch."""
        kwargs = kwargs or {}
        ts = self.timestep_map[t.long()]
        if self.guidance == 'classifier_free':
            y = kwargs['y']
            B = x_t.shape[0]
            out = self.model(torch.cat([x_t, x_t], 0), torch.cat([ts, ts], 0), y=torch.cat([y, torch.zeros_like(y)], 0))
            return out[:B], out[B:]
        return self.model(x_t, ts, **kwargs), None

    def _classifier_grad(self, x_t, t, y):
        """grad_x log p(y | x_t, t) through the user's classifier (diffusion.py:298-304 / 330-336). The classifier
        receives the RESPACED index t, as in the reference."""
        with torch.enable_grad():
            x = x_t.detach().requires_grad_(True)
            log_probs = torch.log_softmax(self.classifier(x, t), dim=-1)
            grabbed = log_probs[range(log_probs.shape[0]), torch.flatten(y)]
            return torch.autograd.grad(grabbed.sum(), x)[0].float().contiguous()

    def _classifier_grad_static(self, x_t, t, y):
        """The same gradient with device-resident arguments only (no host-built index tensors): safe to capture in a CUDA
        graph. `log_probs[range(B), y]` of the reference is a gather along the class axis."""
        with torch.enable_grad():
            x = x_t.detach().clone().requires_grad_(True)
            log_probs = torch.log_softmax(self.classifier(x, t), dim=-1)
            grabbed = log_probs.gather(1, y.reshape(-1, 1))
            return torch.autograd.grad(grabbed.sum(), x)[0].float().contiguous()

    def _step(self, x_t, t, kwargs, clip_x, ddim, noise=None):
        x_t = x_t.contiguous()
        t = t.to(self.device).float().contiguous()
        self._check_step_index(t, 'denoising_step')
        with torch.no_grad():
            out, null = self._model_pair(x_t, t, kwargs)
        grad = None
        if self.guidance == 'classifier':
            grad = self._classifier_grad(x_t, t, kwargs['y'])
        seed, ids = 0, None
        if noise is None:
            if self.noise == 'philox':
                seed = self.seed
                ids = self.sample_ids
            else:
                noise = torch.randn_like(x_t)
        sample = torch.empty_like(x_t)
        pred_x0 = torch.empty_like(x_t)
        with torch.no_grad():
            ops.sampler_step(x_t, out.contiguous(), t, self._tables()['tab'], self._var_code(), ddim=ddim,
                             eta=self.ddim_eta if ddim else 0.0,
                             null_out=null.contiguous() if null is not None else None,
                             cfg_strength=self.strength if null is not None else 0.0,
                             grad=grad, clf_strength=self.strength if grad is not None else 0.0,
                             clip_x0=clip_x, noise=noise, seed=seed, sample_ids=ids, x_prev=sample, pred_x0=pred_x0)
        return sample, pred_x0

    def denoising_step(self, x_t, t, kwargs=None, clip_x=True, noise=None):
        """p(x_{t-1} | x_t): DDPM step with learned / fixed variance (diffusion.py:266-316). Returns
        (sample, pred_x0). `noise` optionally injects the N(0,1) draw."""
        return self._step(x_t, t, kwargs, clip_x, ddim=False, noise=noise)

    def ddim_denoising_step(self, x_t, t, kwargs=None, clip_x=True, noise=None):
        """DDIM step (diffusion.py:318-369). Returns (sample, pred_x0)."""
        return self._step(x_t, t, kwargs, clip_x, ddim=True, noise=noise)

    # ------------------------------------------------------------------------------------------
    # reverse process: the loop
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def denoise(self, x=None, kwargs=None, start_step=None, steps_to_do=None, batch_size=1, ema_params=None,
                progress=True, noise=None):
        """diffusion.py:155-226. Extra keyword `noise`: tensor [n_steps_total, B, C, H, W] whose row t is the
        N(0,1) draw used at respaced step t (parity tests); default follows `self.noise`."""
        if kwargs is None:
            kwargs = {}
        assert ('y' in kwargs.keys() and kwargs['y'] is not None) == self.model.conditional, \
            'pass label iff model is class-conditional'
        if self.model.conditional:
            assert len(kwargs['y']) == batch_size, 'len(labels) != batch size'
        original_params = None
        if ema_params is not None:
            original_params = {}
            for name, param in self.model.named_parameters():
                original_params[name] = param.data
                param.data = ema_params[name].to(self.device)
        try:
            if start_step is None:
                start_step = self.rescaled_num_steps
            if steps_to_do is None or steps_to_do > start_step:
                steps_to_do = start_step
            if x is None:
                assert start_step == self.rescaled_num_steps, 'cannot start from noise with current step that is not T'
                shape = (batch_size, self.model.in_channels, self.model.resolution, self.model.resolution)
                if self.noise == 'philox':
                    ids = self.sample_ids if self.sample_ids is not None else \
                        torch.arange(batch_size, device=self.device, dtype=torch.long)
                    x = ops.philox_normal(shape, self.seed, ids, len(self.betas), self.device)
                else:
                    x = torch.randn(*shape).to(self.device)
            x = x.to(self.device, torch.float32).contiguous()
            if steps_to_do == 0:
                return x
            self._check_step_index(steps_to_do - 1, 'denoise')
            if self.guidance == 'classifier' and not self.classifier_in_graph:
                x = self._denoise_stepwise(x, kwargs, steps_to_do, progress, noise)
            else:
                x = self._denoise_graph(x, kwargs, steps_to_do, progress, noise)
        finally:
            if original_params is not None:
                for name, param in self.model.named_parameters():
                    param.data = original_params[name]
        return x

    def _denoise_stepwise(self, x, kwargs, steps_to_do, progress, noise):
        """Reference-shaped loop over the public step functions (needed for classifier guidance, whose gradient
        goes through the user's torch classifier every step)."""
        indices = reversed(range(steps_to_do))
        if progress:
            indices = tqdm.tqdm(indices, total=steps_to_do)
        for t in indices:
            timestep = torch.full((x.shape[0],), float(t), device=self.device)
            nz = noise[t].to(self.device).contiguous() if noise is not None else None
            step = self.ddim_denoising_step if self.use_ddim else self.denoising_step
            x, _ = step(x, t=timestep, kwargs=kwargs, noise=nz)
        return x

    def _denoise_graph(self, x, kwargs, steps_to_do, progress, noise):
        runner = self.make_step_runner(x, kwargs, steps_to_do=steps_to_do, noise=noise)
        indices = range(steps_to_do)
        if progress:
            indices = tqdm.tqdm(indices, total=steps_to_do)
        for _ in indices:
            runner.step()
        return runner.result()

    def make_step_runner(self, x, kwargs=None, steps_to_do=None, noise=None):
        """Device-resident sampling state: `.step()` advances the whole batch by one denoising step with a single
        CUDA-graph replay; `.result()` returns the current x. Used by `denoise` and by bench.py."""
        if steps_to_do is None:
            steps_to_do = len(self.betas)
        return _StepRunner(self, x.to(self.device, torch.float32).contiguous(), kwargs or {}, steps_to_do, noise)

    # ------------------------------------------------------------------------------------------
    def loss(self, x_0, t, kwargs=None, noise=None):
        """Training loss per sample (diffusion.py:375-406): q-sample x_t (ndb_qsample), differentiable model forward
        (train_forward.py when the model is in training mode), then ONE fused kernel for the simple / KL / KL_rescaled /
        hybrid loss and its gradient w.r.t. the model output (`_FusedLoss`). `t` indexes this object's tables directly, like
        the reference's `extract` -- a t beyond the respaced schedule raises IndexError there and here (the shipped
        trainer draws t in [0, original_num_steps): it needs rescaled_num_steps == original_num_steps)."""
        if kwargs is None:
            kwargs = {}
        x_0 = x_0.to(self.device, torch.float32).contiguous()
        if noise is None:
            noise = torch.randn_like(x_0).to(self.device)
        noise = noise.to(self.device, torch.float32).contiguous()
        t = t.to(self.device)
        self._check_step_index(t, 'loss')
        x_t = self.diffusion_step(x_0=x_0, t=t, noise=noise)
        out = self.model(x_t, self.timestep_map[t.long()], **kwargs)
        return _FusedLoss.apply(out, x_0, x_t, noise, t.float().contiguous(), self._tables()['tab'], self._var_code(),
                                _LOSS_TYPES.index(self.loss_type), int(self.rescaled_num_steps))

    def variational_lower_bound(self, x_0, x_t, t, eps_pred, log_var):
        """diffusion.py:408-438, for callers that already hold eps / log_var (bits per dimension, per sample): the fused
        kernel evaluated as the 'KL' loss on a (eps, log_var) pair presented as a learned-variance model output."""
        out = torch.cat([eps_pred, log_var.expand_as(eps_pred)], dim=1).contiguous().float()
        t = t.to(self.device)
        self._check_step_index(t, 'variational_lower_bound')
        return _FusedLoss.apply(out, x_0.contiguous().float(), x_t.contiguous().float(), torch.zeros_like(x_0),
                                t.float().contiguous(), self._tables()['tab'], L.VAR_LEARNED, L.LOSS_KL,
                                int(self.rescaled_num_steps))


class _FusedLoss(torch.autograd.Function):
    """loss[b] = ndb_diffusion_loss(...); backward: grad_out = (d loss_b / d model_out, written by the same kernel) times the
    upstream gradient of loss[b]."""

    @staticmethod
    def forward(ctx, out, x_0, x_t, noise, t, tables, var_code, loss_code, steps):
        out = out.contiguous()
        need = out.requires_grad
        loss, grad = ops.diffusion_loss(x_0, x_t, noise, out.detach(), t, tables, var_code, loss_code, steps, want_grad=need)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        grad = ctx.grad
        ctx.grad = None
        if grad is None:
            return (None,) * 9
        return (ops.scale_rows_(grad, g_loss.contiguous().float()),) + (None,) * 8


class _StepRunner:
    """One CUDA-graph replay per step: [UNet plan (doubled batch under CFG) -> fused step kernel (writes x_{t-1}
    straight into the UNet's input buffer) -> device-side step counter]. Nothing crosses the host<->device boundary
    inside the loop (the reference does >= 7 H2D copies per step, diffusion.py:216,492)."""

    def __init__(self, diff, x, kwargs, steps_to_do, noise):
        self.diff = diff
        self.B = B = x.shape[0]
        self.cfg = cfg = diff.guidance == 'classifier_free'
        self.clf = diff.guidance == 'classifier'
        self.model_batch = Bm = 2 * B if cfg else B
        model, dev = diff.model, diff.device
        self.eng = eng = model.engine_for(torch.empty((Bm, 0), device=dev))
        # once per chain: also catches in-place writes through param.data, which no version counter sees
        eng.refresh_weights(force=True)
        self.tabs = diff._tables()
        # Loop state and captured graphs live ON THE ENGINE whose buffers they point into (an id(engine) key in a dict
        # of the Diffusion object could be recycled by CPython after the engine is freed, and would replay a graph over
        # freed memory). One state per image batch; one graph per value of everything the capture bakes in.
        st = eng.step_graphs.get(('state', B))
        if st is None:
            st = dict(t_step=torch.zeros((B,), dtype=torch.float32, device=dev), noise=torch.zeros_like(x),
                      ids=torch.arange(B, device=dev, dtype=torch.long), graphs={})
            eng.step_graphs[('state', B)] = st
        self.st = st
        eng.x_in[:B].copy_(x)
        if cfg:
            eng.x_in[B:].copy_(x)
        if model.conditional:
            eng.y_in[:B].copy_(kwargs['y'])
            if cfg:
                eng.y_in[B:].zero_()    # null class 0 (diffusion.py:281)
        self.t = steps_to_do - 1
        st['t_step'].fill_(float(self.t))
        eng.t_in.fill_(float(diff.timestep_map[self.t].item()))
        if diff.sample_ids is not None:
            st['ids'].copy_(diff.sample_ids)
        else:
            st['ids'].copy_(torch.arange(B, device=dev, dtype=torch.long))
        self.noise_rows = noise
        self.use_philox = noise is None and diff.noise == 'philox'
        # every scalar below is a kernel argument frozen at capture time; the reference reads them live each step
        self.gkey = (self.use_philox, int(diff.seed) if self.use_philox else 0, cfg, bool(diff.use_ddim),
                     diff.sampling_var_type, float(diff.strength) if (cfg or self.clf) else 0.0,
                     diff.classifier if self.clf else None,     # the module itself: the key keeps it alive, so no id() reuse
                     float(diff.ddim_eta) if diff.use_ddim else 0.0, self.tabs['tab'].data_ptr(),
                     self.tabs['tmap_f32'].data_ptr())
        self.kernels_per_step = eng.num_kernels + 2

    def _launch(self):
        d, eng, st, B = self.diff, self.eng, self.st, self.B
        eng._launch_all()
        out = eng.out
        grad = None
        if self.clf:
            # d log p(y | x_t, t) / d x_t through the user's classifier: torch autograd, captured with the rest of the step.
            # The classifier sees the RESPACED step index, like the reference's (diffusion.py:301)
            grad = d._classifier_grad_static(eng.x_in[:B], st['t_step'], eng.y_in[:B])
        ops.sampler_step(eng.x_in[:B], out[:B], st['t_step'], self.tabs['tab'], d._var_code(), ddim=d.use_ddim,
                         eta=d.ddim_eta if d.use_ddim else 0.0, null_out=out[B:] if self.cfg else None,
                         cfg_strength=d.strength if self.cfg else 0.0, clip_x0=True,
                         grad=grad, clf_strength=d.strength if self.clf else 0.0,
                         noise=None if self.use_philox else st['noise'], seed=d.seed, sample_ids=st['ids'],
                         x_prev=eng.x_in[:B], x_prev2=eng.x_in[B:] if self.cfg else None)
        ops.advance_step(st['t_step'], eng.t_in, self.tabs['tmap_f32'])

    def step(self):
        d, eng, st = self.diff, self.eng, self.st
        if not self.use_philox:
            if self.noise_rows is not None:
                st['noise'].copy_(self.noise_rows[max(self.t, 0)])
            else:
                st['noise'].normal_()      # torch's generator, like the reference's randn_like per step
        self.t -= 1
        if not d.model.use_cuda_graph:
            self._launch()
            return
        g = st['graphs'].get(self.gkey)
        if g is None:
            # warm up once eagerly (kernel attributes, lazy allocations), restore the state, then capture
            saved = (eng.x_in.clone(), st['t_step'].clone(), eng.t_in.clone())
            self._launch()
            torch.cuda.synchronize(d.device)
            eng.x_in.copy_(saved[0]); st['t_step'].copy_(saved[1]); eng.t_in.copy_(saved[2])
            g = torch.cuda.CUDAGraph()
            with capture_guard(), torch.cuda.graph(g):
                self._launch()
            st['graphs'][self.gkey] = g
        g.replay()
        ops.LAUNCHES += self.kernels_per_step

    def result(self):
        return self.eng.x_in[:self.B].clone()

    def profile_kernels(self):
        """Per-kernel time of one step, measured inside a CUDA graph with %globaltimer stamps between the launches
        (UNetEngine.timeline) plus the two sampler kernels timed the same way. -> {name: {ms, launches, flops, bytes},
        total_ms}. Restores the sampling state."""
        d, eng, st, B = self.diff, self.eng, self.st, self.B
        saved = (eng.x_in.clone(), st['t_step'].clone(), eng.t_in.clone())
        rows, _stamp = eng.timeline()
        # the fused step kernel + the step counter, bracketed the same way
        buf = torch.zeros(4, dtype=torch.int64, device=d.device)

        def tail():
            ops.timestamp(buf, 0)
            ops.timestamp(buf, 1)
            ops.sampler_step(eng.x_in[:B], eng.out[:B], st['t_step'], self.tabs['tab'], d._var_code(), ddim=d.use_ddim,
                             eta=d.ddim_eta if d.use_ddim else 0.0, null_out=eng.out[B:] if self.cfg else None,
                             cfg_strength=d.strength if self.cfg else 0.0, clip_x0=True,
                             noise=None if self.use_philox else st['noise'], seed=d.seed, sample_ids=st['ids'],
                             x_prev=eng.x_in[:B], x_prev2=eng.x_in[B:] if self.cfg else None)
            ops.timestamp(buf, 2)
            ops.advance_step(st['t_step'], eng.t_in, self.tabs['tmap_f32'])
            ops.timestamp(buf, 3)
        tail()
        torch.cuda.synchronize(d.device)
        g = torch.cuda.CUDAGraph()
        with capture_guard(), torch.cuda.graph(g):
            tail()
        g.replay()
        torch.cuda.synchronize(d.device)
        tb = buf.cpu().double()
        cost = float(tb[1] - tb[0])
        px = eng.x_in[:B].numel()
        rows = list(rows) + [('sampler_step', '', max(float(tb[2] - tb[1]) - cost, 0.0) / 1e3, 0, 16 * px),
                             ('advance_step', '', max(float(tb[3] - tb[2]) - cost, 0.0) / 1e3, 0, 0)]
        eng.x_in.copy_(saved[0]); st['t_step'].copy_(saved[1]); eng.t_in.copy_(saved[2])
        out, total = {}, 0.0
        for name, _desc, us, flops, nbytes in rows:
            e = out.setdefault(name, {'ms': 0.0, 'launches': 0, 'flops': 0, 'bytes': 0})
            e['ms'] += us / 1e3
            e['launches'] += 1
            e['flops'] += flops
            e['bytes'] += nbytes
            total += us / 1e3
        out['total_ms'] = total
        return out
