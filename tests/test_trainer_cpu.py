"""CPU: the data-parallel gradient exchange of the training step (trainer.GradBuckets) across two gloo ranks, the loss
oracle against the reference's golden losses, and the differentiable training forward against the reference fixtures.
(The fused loss kernel and the Trainer itself need a GPU: tests/test_train_gpu.py.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _toy():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(12, 33), torch.nn.SiLU(), torch.nn.Linear(33, 20), torch.nn.SiLU(),
                               torch.nn.Linear(20, 5))


def _toy_batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(6, 12, generator=g), torch.randn(6, 5, generator=g)


def _worker(rank, world, port, out):
    from nice_diffusion_b200.trainer import GradBuckets
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _toy()
        # tiny buckets: five of the six parameter tensors get their own collective (reverse registration order)
        buckets = GradBuckets(model.parameters(), bucket_bytes=1024)
        assert len(buckets.buckets) >= 3 and buckets.world == world
        assert all(p.grad is not None and p.grad.data_ptr() >= b['flat'].data_ptr()
                   for b in buckets.buckets for p in b['params'])
        res = []
        for it in range(2):                      # second iteration: hooks re-armed, buffers zeroed
            x, y = _toy_batch(rank + 10 * it)
            ((model(x) - y) ** 2).mean().backward()
            buckets.finish()
            res.append([p.grad.clone() for p in model.parameters()])
            buckets.zero()
            assert all(float(p.grad.abs().max()) == 0.0 for p in model.parameters())
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_grad_buckets_average_across_two_ranks(tmp_path):
    world = 2
    out = str(tmp_path / 'grads.pt')
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out)
    for it in range(2):
        want = None
        for rank in range(world):
            model = _toy()
            x, y = _toy_batch(rank + 10 * it)
            ((model(x) - y) ** 2).mean().backward()
            g = [p.grad.clone() for p in model.parameters()]
            want = g if want is None else [a + b for a, b in zip(want, g)]
        for a, b in zip(got[it], want):
            assert torch.allclose(a, b / world, rtol=1e-6, atol=1e-7)


def test_grad_buckets_single_process_is_a_no_op_exchange():
    from nice_diffusion_b200.trainer import GradBuckets
    model = _toy()
    ref = _toy()
    buckets = GradBuckets(model.parameters(), bucket_bytes=1 << 20)
    assert len(buckets.buckets) == 1 and buckets.world == 1
    x, y = _toy_batch(0)
    ((model(x) - y) ** 2).mean().backward()
    buckets.finish()
    ((ref(x) - y) ** 2).mean().backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        assert torch.equal(p.grad, q.grad)


def test_loss_oracle_matches_reference_goldens():
    """oracle/loss.py on the stored model outputs reproduces the reference's losses bit-for-bit (CPU fp32) and its
    d loss.mean() / d model_out."""
    from oracle import CONFIGS
    from oracle.loss import LossOracle
    from oracle.make_golden import LOSS_CASES, loss_case_inputs
    for name, (cname, dkw) in LOSS_CASES.items():
        g = np.load(os.path.join(GOLD, 'loss_{}.npz'.format(name)))
        inp = loss_case_inputs(name)
        okw = {k: v for k, v in dkw.items() if k != 'loss_type'}
        ora = LossOracle(None, dkw['loss_type'], **okw)
        out = torch.from_numpy(g['out']).requires_grad_(True)
        loss = ora.loss_from_output(inp['x0'], torch.from_numpy(g['x_t']), inp['t'], out, inp['noise'])
        assert torch.equal(loss.detach(), torch.from_numpy(g['loss'])), name
        gout = torch.autograd.grad(loss.mean(), out)[0]
        assert torch.equal(gout, torch.from_numpy(g['g_out'])), name
        assert torch.equal(ora.diffusion_step(inp['x0'], inp['t'], inp['noise']), torch.from_numpy(g['x_t']))


def test_training_forward_reproduces_reference_loss_on_cpu():
    """The differentiable forward (train_forward.py) is plain torch, so it also runs on the CPU: with the golden inputs it
    reproduces the reference's parameter gradients. (The loss kernel is CUDA-only; the oracle's loss stands in here.)"""
    from oracle import CONFIGS
    from oracle.loss import LossOracle
    from oracle.make_golden import LOSS_CASES, loss_case_inputs
    from oracle.weights import random_state_dict
    from nice_diffusion_b200.model import DiffusionModel
    name = 'hybrid_interp'
    cname, dkw = LOSS_CASES[name]
    cfg = dict(CONFIGS[cname]); cfg['dropout'] = 0.0
    model = DiffusionModel(**cfg)
    model.load_state_dict(random_state_dict(cfg, seed=777), strict=True)
    model.train()
    inp = loss_case_inputs(name)
    g = np.load(os.path.join(GOLD, 'loss_{}.npz'.format(name)))
    okw = {k: v for k, v in dkw.items() if k != 'loss_type'}
    ora = LossOracle(lambda xx, tt, y=None: model(xx, tt, y), dkw['loss_type'], **okw)
    loss = ora.loss(inp['x0'], inp['t'], {'y': inp['y']}, inp['noise'])
    assert torch.equal(loss.detach(), torch.from_numpy(g['loss']))
    loss.mean().backward()
    named = dict(model.named_parameters())
    checked = 0
    for key in g.files:
        if key.startswith('grad/'):
            assert torch.equal(named[key[5:]].grad, torch.from_numpy(g[key])), key
            checked += 1
    assert checked >= 3
