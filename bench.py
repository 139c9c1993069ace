#!/usr/bin/env python
"""Benchmark of the sampling hot path (BASELINE.json: images/sec, ImageNet-64 UNet, 250-step respaced DDPM with
learned-interpolated variance, batch 256 sharded over the GPUs of the run: 256 / N images per GPU, strong scaling --
the whole named configuration on one GPU at N = 1, 32 images per GPU at N = 8; `--scaling weak` keeps 32 per GPU at
every N, the round-1 protocol).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config in64|emnist|in128|in64_clf|train_*]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one denoising step (UNet forward + fused sampler step) over the rank's batch, replayed as one CUDA
graph. `value` = images/sec of the whole job for the full chain: N * B / (steps_per_image * t_step), device-timed,
max over ranks. `e2e` = the same metric through the public API (`Diffusion.denoise`) for one complete chain, with
x_T / labels coming from pinned host memory and the finished images copied back, all inside the timed region.
`roofline` = algorithmic FLOPs of every implicit-GEMM conv launch / its duration measured inside the step's CUDA graph
(%globaltimer stamps between the kernels) against the measured dense bf16 peak, plus the whole-forward fraction.
`cpu_baseline` / `--impl reference` = the unmodified reference's --cpu path (baseline/_ref) on the host cores, a
bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (oracle config key, diffusion kwargs, per-GPU batch, steps per image)
    'in64': ('in64', dict(original_num_steps=1000, rescaled_num_steps=250, beta_schedule='cosine',
                          sampling_var_type='learned_interpolation', loss_type='hybrid'), 32, 250),
    'emnist': ('emnist', dict(original_num_steps=1000, rescaled_num_steps=25, beta_schedule='cosine',
                              sampling_var_type='learned_interpolation', loss_type='hybrid',
                              guidance_method='classifier_free', guidance_strength=0.8), 16, 25),
    # BASELINE config 4: classifier-guided ImageNet-64 sampling -- the gradient of an ADM-style noisy classifier (random-init,
    # 41 M parameters; torch autograd, forward + input-backward) is taken every step INSIDE the per-step CUDA graph
    'in64_clf': ('in64', dict(original_num_steps=1000, rescaled_num_steps=250, beta_schedule='cosine',
                              sampling_var_type='learned_interpolation', loss_type='hybrid',
                              guidance_method='classifier', guidance_strength=1.0), 32, 250),
    'in128': ('in128', dict(original_num_steps=1000, rescaled_num_steps=50, beta_schedule='linear',
                            sampling_var_type='learned_interpolation', loss_type='hybrid', use_ddim=True, ddim_eta=0.0,
                            guidance_method='classifier_free', guidance_strength=0.8), 8, 50),
}
# algorithmic FLOPs per image per UNet forward (counted on the reference modules)
FLOPS_PER_FWD = {'in64': 219356037120, 'emnist': 3998584832, 'in128': 613855789056, 'in64_clf': 219356037120}


# BASELINE.json configs[1]: "batch 256 sharded over 8 x B200"
IN64_GLOBAL_BATCH = 256


def per_gpu_batch(config, world, batch=None, scaling='strong'):
    """-> (per-GPU batch, value of the line's `scaling` key). ImageNet-64 (the configuration the metric is quoted on) runs
    the batch the configuration names on however many GPUs the run has: total work fixed, 256 / N images per GPU. An explicit
    --batch, --scaling weak and the other workloads (single-GPU configurations) keep a fixed per-GPU batch."""
    if batch:
        return int(batch), 'weak'
    if config == 'in64' and scaling == 'strong':
        if IN64_GLOBAL_BATCH % world != 0:
            raise SystemExit('bench.py: {} GPUs do not divide the global batch of {}'.format(world, IN64_GLOBAL_BATCH))
        return IN64_GLOBAL_BATCH // world, 'strong'
    table = WORKLOADS if config in WORKLOADS else TRAIN_WORKLOADS
    return (table[config][2] if config in WORKLOADS else table[config][1]), 'weak'


def conv_traffic_capture(config, batch, profiles_dir=None):
    """The committed ncu capture (profiles/*_conv_traffic*.json: DRAM bytes of every conv GEMM launch of one step) that
    matches this run: ImageNet-64 at the same per-GPU batch (files without the key were taken at 32), newest tag last.
    None when there is none -- `roofline.traffic` is then null rather than a number from another workload."""
    import glob
    if config != 'in64':
        return None
    d = profiles_dir or os.path.join(ROOT, 'profiles')
    best = None
    for path in sorted(glob.glob(os.path.join(d, '*_conv_traffic*.json'))):
        try:
            tr = json.load(open(path))
        except (OSError, ValueError):
            continue
        if tr.get('per_gpu_batch', WORKLOADS['in64'][2]) == batch and 'dram_bytes_per_launch_avg' in tr:
            best = dict(tr, file=os.path.basename(path))
    return best


def model_config(name):
    from nice_diffusion_b200 import default_args as da
    cfg = dict({'in64': da.OPENAI_64_MODEL_ARGS, 'emnist': da.EMNIST_MODEL_ARGS, 'in128': da.OPENAI_128_MODEL_ARGS,
                'in64_clf': da.OPENAI_64_MODEL_ARGS}[name])
    if name == 'in128':
        cfg['num_classes'] = 1001  # CFG adds the null class (utils.py:211-212)
    return cfg


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    hbm=p['hbm_gbs'], source='measured')
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def import_reference():
    """The UNMODIFIED reference package, pip-installed into baseline/_ref (git-ignored; it travels to the GPU box with
    the snapshot). nicediffusion/utils.py imports matplotlib at module top, which this image lacks: a stub module
    satisfies the import (nothing on the sampling path touches it). Returns (model module, diffusion module) or None."""
    import types
    ref_dir = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'nicediffusion')):
        return None
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        for name in ('matplotlib', 'matplotlib.pyplot'):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import nicediffusion.model as rm
        import nicediffusion.diffusion as rd
        return rm, rd
    except Exception as e:  # noqa: BLE001
        print('bench.py: reference import failed: {}'.format(e), file=sys.stderr)
        return None


def host_cores():
    # all host cores this process may use (torchrun pins OMP_NUM_THREADS=1 in its children; undo that for the CPU arm)
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_arm(workload, steps, warmup, sample_batch=None):
    """The reference's own --cpu path: the real `nicediffusion.Diffusion.denoise(device=cpu)` from baseline/_ref
    (kind 'reference') -- or, if that install is missing, the oracle port of it (kind 'port') -- for `steps` denoising
    steps on a bounded sample of the workload's batch."""
    ref = import_reference()
    if ref is not None:
        return real_reference_arm(ref, workload, steps, warmup, sample_batch)
    return oracle_port_arm(workload, steps, warmup, sample_batch)


def real_reference_arm(ref, workload, steps, warmup, sample_batch):
    rm, rd = ref
    from nice_diffusion_b200.utils import randomize_zero_init_
    cname, dkw, _, per_image = WORKLOADS[workload]
    cfg = model_config(workload)
    if sample_batch is None:
        sample_batch = {'in64': 2, 'emnist': 16, 'in128': 1, 'in64_clf': 2}[workload]
    torch.set_num_threads(max(1, host_cores()))
    torch.manual_seed(0)
    model = rm.DiffusionModel(**cfg)
    randomize_zero_init_(model, seed=0)
    cpu = torch.device('cpu')
    if dkw.get('guidance_method') == 'classifier':
        # the reference ships no classifier; its API takes any torch module -- the same NoisyClassifier the GPU arm uses
        from nice_diffusion_b200.classifier import NoisyClassifier
        dkw = dict(dkw, classifier=NoisyClassifier(64, 3, 128, 1000, 2, (32, 16, 8), channel_mult=(1, 2, 3, 4), num_head_channels=64,
                                                   steps_scale=dkw['original_num_steps'] / dkw['rescaled_num_steps']))
    diff = rd.Diffusion(model=model, device=cpu, **dkw)
    x = torch.randn(sample_batch, cfg['in_channels'], cfg['resolution'], cfg['resolution'])
    y = torch.randint(1, cfg['num_classes'], (sample_batch,))
    # `denoise(x, steps_to_do=k)` runs the k steps t = k-1 .. 0 (diffusion.py:192-219); a step's cost does not depend on t
    diff.denoise(x=x.clone(), kwargs={'y': y}, batch_size=sample_batch, progress=False, steps_to_do=max(1, warmup))
    t0 = time.perf_counter()
    diff.denoise(x=x.clone(), kwargs={'y': y}, batch_size=sample_batch, progress=False, steps_to_do=steps)
    t_step = (time.perf_counter() - t0) / steps
    return dict(value=sample_batch / (per_image * t_step), ms_per_step=1e3 * t_step, cores=torch.get_num_threads(),
                kind='reference', batch=sample_batch,
                sample='nicediffusion.Diffusion.denoise(device=cpu) from baseline/_ref (unmodified reference, fp32): {} '
                       'denoising steps (+{} warm-up) of the {} chain at batch {}; images/sec = batch / ({} steps x time '
                       'per step)'.format(steps, warmup, workload, sample_batch, per_image))


def gpu_eager_reference(workload, B, dev, steps=3):
    """Informational: the unmodified reference run eagerly on the SAME B200 -- torch's default
    precision (cuDNN convs in TF32) and under bf16 autocast -- for a few denoising steps at this arm's batch."""
    ref = import_reference()
    if ref is None:
        return None
    rm, rd = ref
    from nice_diffusion_b200.utils import randomize_zero_init_
    cname, dkw, _, per_image = WORKLOADS[workload]
    cfg = model_config(workload)
    out = {'batch': B, 'steps_timed': steps, 'unit': 'images/sec',
           'what': 'nicediffusion (baseline/_ref) Diffusion.denoise on cuda, eager PyTorch, same batch; extrapolated '
                   'from {} steps like every other number here'.format(steps)}
    try:
        torch.manual_seed(0)
        model = rm.DiffusionModel(**cfg)
        randomize_zero_init_(model, seed=0)
        diff = rd.Diffusion(model=model, device=dev, **dkw)
        x = torch.randn(B, cfg['in_channels'], cfg['resolution'], cfg['resolution'], device=dev)
        y = torch.randint(1, cfg['num_classes'], (B,), device=dev)
        for key, ctx in (('tf32_default', None), ('bf16_autocast', torch.bfloat16)):
            def run(k):
                if ctx is None:
                    return diff.denoise(x=x.clone(), kwargs={'y': y}, batch_size=B, progress=False, steps_to_do=k)
                with torch.autocast('cuda', dtype=ctx):
                    return diff.denoise(x=x.clone(), kwargs={'y': y}, batch_size=B, progress=False, steps_to_do=k)
            run(2)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(steps)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[key] = {'ms_per_step': ms, 'value': B / (per_image * ms * 1e-3)}
        del model, diff
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['error'] = repr(e)[:200]
    return out


def oracle_port_arm(workload, steps, warmup, sample_batch=None):
    """Fallback when baseline/_ref is absent: the oracle port of the reference's --cpu path."""
    from oracle import CONFIGS
    from oracle.sampler import Sampler
    from oracle.unet import unet_forward
    from oracle.weights import random_state_dict
    import numpy as np
    cname, dkw, _, per_image = WORKLOADS[workload]
    cfg = CONFIGS[cname]
    if workload == 'in128':
        cfg = dict(cfg); cfg['num_classes'] = 1001
    if sample_batch is None:
        sample_batch = {'in64': 2, 'emnist': 16, 'in128': 1, 'in64_clf': 2}[workload]
    torch.set_num_threads(max(1, host_cores()))
    sd = random_state_dict(cfg, seed=0)
    okw = {k: v for k, v in dkw.items() if k != 'loss_type'}

    def model_fn(x, t, y=None):
        return unet_forward(cfg, sd, x, t, y)
    s = Sampler(model_fn, **okw)
    rs = np.random.RandomState(0)
    x = torch.from_numpy(rs.standard_normal((sample_batch, cfg['in_channels'], cfg['resolution'],
                                             cfg['resolution'])).astype(np.float32))
    y = torch.from_numpy(rs.randint(1, cfg['num_classes'], size=(sample_batch,)).astype(np.int64))
    R = len(s.tab.betas)
    t_hi = R - 1
    step = s.ddim_step if okw.get('use_ddim') else s.ddpm_step
    with torch.no_grad():
        for i in range(warmup):
            x, _ = step(x, torch.full((sample_batch,), float(t_hi - i)), {'y': y}, torch.randn_like(x))
        t0 = time.perf_counter()
        for i in range(steps):
            x, _ = step(x, torch.full((sample_batch,), float(max(t_hi - warmup - i, 0))), {'y': y}, torch.randn_like(x))
        dt = time.perf_counter() - t0
    t_step = dt / steps
    return dict(value=sample_batch / (per_image * t_step), ms_per_step=1e3 * t_step, cores=torch.get_num_threads(),
                kind='port',
                sample='{} denoising steps (+{} warm-up) of the {} chain at batch {} on the host, oracle port of the '
                       'reference --cpu path; images/sec = batch / ({} steps x time per step)'.format(
                           steps, warmup, workload, sample_batch, per_image), batch=sample_batch)


TRAIN_WORKLOADS = {
    # BASELINE config 5: hybrid-loss training step, synthetic batches, data-parallel gradient all-reduce over NCCL
    # name: (model preset, per-GPU batch); T = R = 1000 (the setting the reference's trainer needs)
    'train_emnist': ('emnist', 468),      # batch size of the reference's scripts/train.py:25
    'train_in64': ('in64', 32),
}


def train_arm(args, rank, world, local_rank):
    """One training step = q-sample + differentiable forward + fused loss kernel + backward (bucketed NCCL all-reduce
    overlapped with it) + AdamW + EMA, through nice_diffusion_b200.trainer.Trainer on synthetic host batches."""
    import numpy as np
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from nice_diffusion_b200 import ops
    from nice_diffusion_b200.model import DiffusionModel
    from nice_diffusion_b200.train_forward import count_flops
    from nice_diffusion_b200.trainer import Trainer
    from nice_diffusion_b200.utils import randomize_zero_init_
    preset, B = TRAIN_WORKLOADS[args.config]
    if args.batch:
        B = args.batch
    cfg = model_config(preset)
    torch.manual_seed(0)
    np.random.seed(0)
    model = randomize_zero_init_(DiffusionModel(**cfg), seed=0)
    dargs = dict(original_num_steps=1000, rescaled_num_steps=1000, beta_schedule='cosine',
                 sampling_var_type='learned_interpolation', loss_type='hybrid', guidance_method='classifier_free',
                 guidance_strength=0.8)
    trainer = Trainer(model=model, diffusion_args=dargs, dataloader=None, iterations=0, batch_size=B, lr=1e-4,
                      weight_decay=0.001, device=dev, ema_clone=True)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = (torch.rand(B, cfg['in_channels'], cfg['resolution'], cfg['resolution'], generator=g) * 2 - 1).pin_memory()
    y_host = torch.randint(1, cfg['num_classes'], (B,), generator=g).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(i):
        x = x_host.to(dev, non_blocking=True)
        y = y_host.to(dev, non_blocking=True)
        return trainer.train_step(x, y, step=i)
    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        loss = step(i)
    loss_host = float(loss.item())                 # the step's result crosses back to the host inside the timed region
    e1.record()
    barrier()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    # the gradient exchange alone (same buckets, nothing to overlap with): achieved bus bandwidth
    ar = None
    if dist is not None:
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a0.record()
        for _ in range(reps):
            works = [dist.all_reduce(b['flat'], async_op=True) for b in trainer.buckets.buckets]
            for w in works:
                w.wait()
        a1.record()
        barrier()
        ar_ms = a0.elapsed_time(a1) / reps
        nbytes = trainer.buckets.bytes
        ar = {'bytes': nbytes, 'buckets': len(trainer.buckets.buckets), 'ms': ar_ms,
              'bus_gbs': 2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
              'what': 'all-reduce of the fp32 gradient buckets alone; bus bandwidth = 2 (N-1)/N x bytes / time'}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    fwd = count_flops(model, B)
    tflops = 3.0 * fwd / (ms_per_step * 1e-3) / 1e12
    value = world * B / (ms_per_step * 1e-3)
    line = {'metric': 'images/sec (hybrid-loss training step, {} UNet)'.format(preset), 'value': value, 'unit': 'images/sec',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (torch defaults: cuDNN convolutions may use TF32, matmuls fp32)', 'data': 'synthetic',
            'config': {'workload': '{}: hybrid loss, T = R = 1000, cosine schedule, AdamW, EMA, per-GPU batch {} (global {})'.format(
                args.config, B, B * world), 'per_gpu_batch': B, 'global_batch': B * world,
                'parallelism': 'data parallel x{}: replicas, bucketed NCCL all-reduce of {} MB of fp32 gradients overlapped '
                               'with backward'.format(world, trainer.buckets.bytes >> 20),
                'forward': 'differentiable torch forward over the model parameters (train_forward.py) + fused ndb200 loss kernel',
                'l2': 'weights + activations >> L2'},
            'clocks': clk,
            'e2e': {'value': value, 'unit': 'images/sec', 'h2d_bytes_per_step': x_host.numel() * 4 + y_host.numel() * 8,
                    'd2h_bytes_per_step': 4, 'what': 'the timed loop IS the public API (Trainer.train_step) fed from pinned host '
                                                     'batches; the last loss is read back inside the timed region'},
            'gpu_launches': ops.LAUNCHES - launches0,
            'roofline': {'bound': 'tensor', 'achieved': tflops, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
                         'frac': tflops / pk['bf16_sustained'], 'traffic': None,
                         'what': '3 x forward FLOPs (forward + input- and weight-gradient GEMMs) / step time; fp32 torch '
                                 'kernels, so the bf16 tensor peak is an upper bound they cannot reach'},
            'allreduce': ar, 'loss': loss_host, 'cpu_baseline': None}
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE line (the JSON result): everything else any library prints there (e.g. NCCL's version banner
    under NCCL_DEBUG) is sent to stderr by pointing fd 1 at fd 2 for the rest of the process."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', type=str, default='ndb200', choices=['ndb200', 'reference'])
    ap.add_argument('--config', type=str, default='in64', choices=sorted(WORKLOADS) + sorted(TRAIN_WORKLOADS))
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch, fixed for every N (default: see --scaling)')
    ap.add_argument('--scaling', type=str, default='strong', choices=['strong', 'weak'],
                    help="in64 only. strong (default): the batch BASELINE config 2 names -- 256 images -- on the GPUs of the "
                         "run, 256 / N per GPU (N=1: the whole configuration on one GPU; N=8: 32 per GPU as the config says). "
                         "weak: 32 images per GPU at every N (the round-1 protocol)")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--eager', action='store_true', help='launch kernels one by one instead of replaying CUDA graphs (profiling)')
    ap.add_argument('--operands', type=str, default='fp16', choices=['fp16', 'fp32x'],
                    help="tensor-core operand mode: fp16 (default, 2e-3 class) or fp32x = fp32-class (three fp16 products per "
                         "multiply, chunked accumulation; 1e-5 class)")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-reference', action='store_true', help='skip the informational eager-reference-on-GPU leg')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.config in TRAIN_WORKLOADS:
        if args.impl == 'reference':
            if rank == 0:
                _emit({'impl': 'reference', 'unavailable': 'the training configs have no CPU reference arm (use --config in64)'})
            return
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device')
        return train_arm(args, rank, world, local_rank)
    cname, dkw, B, per_image = WORKLOADS[args.config]
    B, scaling = per_gpu_batch(args.config, world, args.batch, args.scaling)
    metric = 'images/sec ({} UNet, {}-step {} sampling)'.format(
        {'in64': 'ImageNet-64', 'emnist': 'EMNIST', 'in128': 'ImageNet-128', 'in64_clf': 'ImageNet-64 + noisy-classifier guidance'}[args.config], per_image,
        'DDIM' if dkw.get('use_ddim') else 'DDPM')
    config = {'workload': '{}: {} steps/image, {}, per-GPU batch {} (global {}{}), {}'.format(
        args.config, per_image, dkw['sampling_var_type'], B, B * world,
        ': the batch BASELINE config 2 names, sharded over the GPUs of the run' if scaling == 'strong' else '',
        'classifier-free guidance as a doubled batch' if dkw.get('guidance_method') else 'no guidance'),
        'global_batch': B * world, 'per_gpu_batch': B, 'steps_per_image': per_image,
        'parallelism': 'batch-sharded x{} (no collective on the data path)'.format(world),
        'operands': 'fp16 tensor-core operands, fp32 accumulate, fp32 residual stream',
        'l2': 'no flush needed: one step touches ~0.6 GB of weights + several GB of activations (>> 126 MB L2)'}

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == 'reference':
        if rank != 0:
            return
        small = args.steps > 40
        r = cpu_reference_arm(args.config, args.steps, args.warmup,
                              sample_batch=({'in64': 1, 'emnist': 16, 'in128': 1, 'in64_clf': 1}[args.config] if small else None))
        # this arm's TRUE batch: a bounded sample of the workload on the host (one process, no GPUs)
        config = dict(config, workload=config['workload'] + ' -- CPU arm: bounded sample at batch {} on {} host threads'.format(
            r['batch'], r['cores']), per_gpu_batch=r['batch'], global_batch=r['batch'],
            parallelism='one host process, {} threads'.format(r['cores']), operands='fp32 (PyTorch CPU kernels)', l2='n/a')
        line = {'impl': 'reference', 'metric': metric, 'value': r['value'], 'unit': 'images/sec', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
                'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
                'cpu_baseline': {'value': r['value'], 'unit': 'images/sec', 'cores': r['cores'], 'kind': r['kind'],
                                 'sample': r['sample']},
                'e2e': {'value': r['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        _emit(line)
        return

    # ------------------------------------------------------------------ ndb200 arm
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the ndb200 path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    from nice_diffusion_b200 import ops
    from nice_diffusion_b200.diffusion import Diffusion
    from nice_diffusion_b200.model import DiffusionModel
    from nice_diffusion_b200.utils import randomize_zero_init_

    torch.manual_seed(0)
    model = DiffusionModel(**model_config(args.config))
    randomize_zero_init_(model, seed=0)
    model = model.to(dev).eval()
    if args.operands == 'fp32x':
        from nice_diffusion_b200 import _lib as _L
        model.operand_format = _L.FMT_FP32X
        config['operands'] = 'fp32-class: fp16 (hi, lo) operand pairs, three tensor-core products per multiply, K accumulated ' \
                             'in 128-element chunks added in fp32 round-to-nearest, fp32 residual stream'
    model.use_cuda_graph = not args.eager
    if dkw.get('guidance_method') == 'classifier':
        from nice_diffusion_b200.classifier import NoisyClassifier
        dkw = dict(dkw, classifier=NoisyClassifier(64, 3, 128, 1000, 2, (32, 16, 8), channel_mult=(1, 2, 3, 4), num_head_channels=64,
                                                   steps_scale=dkw['original_num_steps'] / dkw['rescaled_num_steps']))
        config['classifier'] = 'ADM-style encoder + attention pooling, 41 M parameters, fp32 torch autograd inside the step graph'
    diff = Diffusion(model=model, device=dev, **dkw)
    diff.noise, diff.seed = 'philox', 1234
    ids = torch.arange(rank * B, (rank + 1) * B, device=dev, dtype=torch.long)   # global sample indices
    diff.sample_ids = ids
    g = torch.Generator(device='cpu').manual_seed(1)
    labels_host = torch.randint(1, 1000 if args.config != 'emnist' else 27, (B * world,), generator=g)[rank * B:(rank + 1) * B]
    labels_host = labels_host.pin_memory()
    x_host = torch.randn((B, model.in_channels, model.resolution, model.resolution), generator=g).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- kernel-only timing: K graph replays of one denoising step, inputs resident in HBM ----
    runner = diff.make_step_runner(x_host.to(dev), {'y': labels_host.to(dev)})
    if args.eager:
        runner.eng.serial = True          # profiling runs: one stream, kernels do not overlap
    for _ in range(args.warmup):
        runner.step()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.LAUNCHES
    if os.environ.get('NDB_CUDA_PROFILER'):
        torch.cuda.profiler.start()       # ncu --profile-from-start off: capture exactly the timed steps
    ev0.record()
    for _ in range(args.steps):
        runner.step()
    ev1.record()
    if os.environ.get('NDB_CUDA_PROFILER'):
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
    barrier()
    launches = ops.LAUNCHES - launches0
    clk = clocks.stop()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tmax = torch.tensor([ms], device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = world * B / (per_image * ms_per_step * 1e-3)

    # ---- end to end through the public API: one complete chain, host buffers in and out ----
    e2e = None
    if not args.no_e2e:
        diff.noise = 'philox'
        out_host = torch.empty((B, model.in_channels, model.resolution, model.resolution)).pin_memory()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x_dev = x_host.to(dev, non_blocking=True)
        y_dev = labels_host.to(dev, non_blocking=True)
        out = diff.denoise(x=x_dev, kwargs={'y': y_dev}, batch_size=B, progress=False)
        out_host.copy_(out, non_blocking=True)
        e1.record()
        barrier()
        if not bool(torch.isfinite(out_host).all()):
            raise SystemExit('bench.py: the end-to-end chain produced non-finite images')
        ems = e0.elapsed_time(e1)
        if dist is not None:
            tmax = torch.tensor([ems], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ems = float(tmax.item())
        h2d = x_host.numel() * 4 + labels_host.numel() * 8
        d2h = out_host.numel() * 4
        e2e = {'value': world * B / (ems * 1e-3), 'unit': 'images/sec', 'h2d_bytes_per_step': h2d / per_image,
               'd2h_bytes_per_step': d2h / per_image, 'steps_per_call': per_image, 'h2d_bytes_per_call': h2d,
               'd2h_bytes_per_call': d2h, 'ms_per_call': ems,
               'what': 'Diffusion.denoise(x=pinned host x_T -> device, labels) for the full chain + D2H of the images'}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline: every launch of the step timed IN the step's graph (globaltimer stamps between kernels, warm caches,
    #      no host launch latency: UNetEngine.timeline), aggregated per kernel ----
    pk = peaks()
    prof = runner.profile_kernels()
    conv = prof['conv_gemm']
    achieved = conv['flops'] / (conv['ms'] * 1e-3) / 1e12
    fwd_tflops = FLOPS_PER_FWD[args.config] * runner.model_batch / (ms_per_step * 1e-3) / 1e12
    gn = prof.get('gn_apply', {'ms': 0.0, 'bytes': 0, 'launches': 0})
    at = prof.get('attention', {'ms': 0.0, 'flops': 0, 'launches': 0})
    roofline = {'bound': 'tensor', 'kernel': 'conv_gemm_kernel (tcgen05 implicit GEMM, all conv / linear layers)',
                'achieved': achieved, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s', 'frac': achieved / pk['bf16_sustained'],
                'peak_kind': pk['source'] + ' sustained dense bf16 (kernel timed inside a long step)',
                'frac_of_burst_peak': achieved / pk['bf16'], 'traffic': None,
                'launches_per_step': conv['launches'], 'ms_per_step_in_kernel': conv['ms'],
                'share_of_step': conv['ms'] / prof['total_ms'],
                # the whole UNet forward (the number the north-star target is written against): algorithmic FLOPs of one
                # forward x batch / graph-replayed step time
                'step_tflops': fwd_tflops, 'step_frac': fwd_tflops / pk['bf16_sustained'],
                'step_frac_of_burst': fwd_tflops / pk['bf16'],
                # HBM-bound norm kernel: algorithmic bytes (input once + every output once) / its time in the step
                'gn_apply_gbs': gn['bytes'] / (gn['ms'] * 1e-3) / 1e9 if gn['ms'] > 0 else None,
                'gn_apply_frac': gn['bytes'] / (gn['ms'] * 1e-3) / 1e9 / pk['hbm'] if gn['ms'] > 0 else None,
                'gn_apply_ms': gn['ms'], 'gn_apply_launches': gn['launches'], 'hbm_peak_gbs': pk['hbm'],
                'attention_tflops': at['flops'] / (at['ms'] * 1e-3) / 1e12 if at['ms'] > 0 else None,
                'attention_ms': at['ms'], 'attention_launches': at['launches'],
                'timing': 'per-kernel times: %globaltimer stamps between the launches inside one CUDA graph of the step',
                'other_kernels_ms': {k: v['ms'] for k, v in prof.items() if isinstance(v, dict) and k != 'conv_gemm'}}

    # informational, N = 1 of the ImageNet-64 line only: the same step at the per-GPU batch of the 8-GPU run (32 images), so
    # that one line carries both ends of the "batch 256 sharded over 8 GPUs" configuration
    shard = WORKLOADS['in64'][2]
    if args.config == 'in64' and world == 1 and B > shard and not args.eager:
        try:
            diff.sample_ids = ids[:shard]
            r32 = diff.make_step_runner(x_host[:shard].to(dev), {'y': labels_host[:shard].to(dev)})
            for _ in range(max(args.warmup, 3)):
                r32.step()
            torch.cuda.synchronize(dev)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(20):
                r32.step()
            s1.record()
            torch.cuda.synchronize(dev)
            ms32 = s0.elapsed_time(s1) / 20
            tf32 = FLOPS_PER_FWD[args.config] * r32.model_batch / (ms32 * 1e-3) / 1e12
            roofline['shard_of_8_gpu_run'] = {'per_gpu_batch': shard, 'steps': 20, 'ms_per_step': ms32,
                                              'images_per_sec_per_gpu': shard / (per_image * ms32 * 1e-3),
                                              'step_tflops': tf32, 'step_frac': tf32 / pk['bf16_sustained']}
            del r32
        except Exception as e:  # noqa: BLE001  (information only: never lose the line over it)
            roofline['shard_of_8_gpu_run'] = {'error': repr(e)[:200]}
        finally:
            diff.sample_ids = ids

    # DRAM traffic of the dominant kernel: from the committed ncu pass of the same step (profiles/), not measured live
    tr = conv_traffic_capture(args.config, B)
    if tr is not None:
        roofline['traffic'] = tr['dram_bytes_per_launch_avg']
        roofline['traffic_unit'] = 'bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum averaged over the %d ' \
                                   'conv_gemm launches of one step; %s)' % (tr['launches'], tr['file'])

    gpu_ref = None
    if not args.no_gpu_reference and world == 1:
        del runner
        model.release_engines()
        torch.cuda.empty_cache()
        gpu_ref = gpu_eager_reference(args.config, B, dev)

    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_reference_arm(args.config, 3, 1)
        cpu = {'value': r['value'], 'unit': 'images/sec', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}

    line = {'metric': metric, 'value': value, 'unit': 'images/sec', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': scaling,
            'vs_baseline': None, 'dtype': 'fp16' if args.operands == 'fp16' else 'fp32x', 'data': 'synthetic', 'config': config,
            'clocks': clk,
            'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
            'gpu_eager_reference': gpu_ref}
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
