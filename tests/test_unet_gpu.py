"""UNet forward parity on the GPU: product DiffusionModel (ndb200 kernels) vs the oracle (CPU fp32 restatement,
pinned bit-exact to the reference) and vs the committed golden outputs of the reference itself.

Metric: max|out - ref| / max|ref| over the whole output tensor -- element-wise
relative error is meaningless because eps/variance outputs cross zero. Bar from BASELINE.json: 2e-3 in the 16-bit
operand mode. Operands are fp16 (not bf16): same tensor-core rate, 3 more mantissa bits; bf16 operands measure
6.7e-3 on these weights (oracle/precision_study.py) and cannot meet the bar.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 2e-3


def build(cfg_name, seed, dev):
    from oracle import CONFIGS
    from oracle.weights import random_state_dict
    from nice_diffusion_b200.model import DiffusionModel
    cfg = CONFIGS[cfg_name]
    sd = random_state_dict(cfg, seed=seed)
    model = DiffusionModel(**cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.to(dev).eval()


def norm_err(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize('name', ['tiny', 'tiny_legacy', 'tiny_pool', 'tiny96', 'emnist', 'in64', 'in128'])
def test_forward_matches_reference_golden(cuda_dev, name):
    cfg, sd, model = build(name, 1234, cuda_dev)
    g = np.load(os.path.join(GOLD, 'unet_{}.npz'.format(name)))
    x = torch.from_numpy(g['x']).to(cuda_dev)
    t = torch.from_numpy(g['t']).to(cuda_dev)
    y = torch.from_numpy(g['y']).to(cuda_dev) if 'y' in g.files else None
    with torch.no_grad():
        out = model(x, t, y) if y is not None else model(x, t)
    ref = torch.from_numpy(g['out'])
    assert out.shape == ref.shape and out.dtype == torch.float32
    err = norm_err(out.cpu(), ref)
    print(name, 'max|d|/max|ref| =', err)
    assert err < TOL
    # eps and variance halves separately (BASELINE: "eps and variance outputs")
    if ref.shape[1] % 2 == 0:
        h = ref.shape[1] // 2
        assert norm_err(out.cpu()[:, :h], ref[:, :h]) < TOL
        assert norm_err(out.cpu()[:, h:], ref[:, h:]) < TOL


@pytest.mark.parametrize('name,batch', [('tiny', 5), ('emnist', 3)])
def test_forward_matches_oracle_other_batches(cuda_dev, name, batch):
    """Fresh inputs / ragged batch sizes against the oracle run live on the CPU; also exercises graph replay
    (second call) and eager mode agreeing bit-for-bit."""
    from oracle.unet import unet_forward
    from oracle.make_golden import make_inputs
    cfg, sd, model = build(name, 77, cuda_dev)
    x, t, y = make_inputs(cfg, batch, seed=5)
    ref = unet_forward(cfg, sd, x, t, y)
    d = cuda_dev
    out1 = model(x.to(d), t.to(d), y.to(d))
    out2 = model(x.to(d), t.to(d), y.to(d))          # graph replay
    assert torch.equal(out1, out2)
    assert norm_err(out1.cpu(), ref) < TOL
    model.use_cuda_graph = False
    model.release_engines()
    out3 = model(x.to(d), t.to(d), y.to(d))          # eager launches
    assert torch.equal(out1, out3)


def test_weight_refresh_and_errors(cuda_dev):
    from oracle.unet import unet_forward
    from oracle.weights import random_state_dict
    from oracle.make_golden import make_inputs
    from nice_diffusion_b200 import _lib
    cfg, sd, model = build('tiny', 1, cuda_dev)
    x, t, y = make_inputs(cfg, 2, seed=5)
    d = cuda_dev
    a = model(x.to(d), t.to(d), y.to(d))
    # load_state_dict and the EMA-style param.data swap (diffusion.py:185-189) must invalidate the packed weights
    sd2 = random_state_dict(cfg, seed=2)
    model.load_state_dict(sd2)
    b = model(x.to(d), t.to(d), y.to(d))
    assert norm_err(b.cpu(), unet_forward(cfg, sd2, x, t, y)) < TOL
    saved = {n: p.data for n, p in model.named_parameters()}
    for n, p in model.named_parameters():
        p.data = sd[n].to(d)
    c = model(x.to(d), t.to(d), y.to(d))
    assert torch.equal(a, c)
    for n, p in model.named_parameters():
        p.data = saved[n]
    assert torch.equal(model(x.to(d), t.to(d), y.to(d)), b)
    # reference assertions
    with pytest.raises(AssertionError):
        model(x.to(d), t.to(d))                       # conditional model without y
    with pytest.raises(AssertionError):
        model(x.to(d)[:, :, :8, :8], t.to(d), y.to(d))  # wrong resolution
    # no CPU fallback
    with pytest.raises(_lib.NdbError):
        model(x, t, y)


def test_full_size_batch_properties_in64(cuda_dev):
    """BASELINE size (ImageNet-64 UNet, 32 images per GPU), checked through properties that do not need the CPU oracle:
    * the golden sample of the reference, embedded at two positions of a full batch, still matches the reference output;
    * a sample's output does not depend on its batch neighbours or position (per-sample GroupNorm / attention): rows of
      the B=32 run equal the same samples run as a batch of 2 up to operand-rounding noise (a different tile -> CTA
      assignment changes the fp32 summation order of the GroupNorm statistics by ~1e-7, which flips the fp16 rounding of a
      few operands downstream: same order as the 16-bit error budget, far below a real cross-sample leak, which is O(1));
    * re-running the full batch (graph replay) is bit-identical."""
    cfg, sd, model = build('in64', 1234, cuda_dev)
    g = np.load(os.path.join(GOLD, 'unet_in64.npz'))
    gx, gt, gy = (torch.from_numpy(g[k]) for k in ('x', 't', 'y'))
    ref = torch.from_numpy(g['out'])
    B = 32
    gen = torch.Generator(device='cpu').manual_seed(99)
    x = torch.randn(B, 3, 64, 64, generator=gen)
    t = torch.tensor([2.0 + 4 * (i % 250) for i in range(B)])          # entries of the R=250 timestep map
    y = torch.randint(0, 1000, (B,), generator=gen)
    n = gx.shape[0]
    for pos in (0, B - n):
        x[pos:pos + n], t[pos:pos + n], y[pos:pos + n] = gx, gt.float(), gy
    d = cuda_dev
    with torch.no_grad():
        full = model(x.to(d), t.to(d), y.to(d))
        again = model(x.to(d), t.to(d), y.to(d))
        pair = model(x[5:7].to(d), t[5:7].to(d), y[5:7].to(d))
    assert torch.equal(full, again)
    for pos in (0, B - n):
        assert norm_err(full[pos:pos + n].cpu(), ref) < TOL
    assert norm_err(full[5:7].cpu(), pair.cpu()) < 1.5e-3
    assert torch.isfinite(full).all()


# ----------------------------------------------------------------------------------------------------------------------
# fp32-class mode (BASELINE north_star: "1e-5 in fp32 mode")
# ----------------------------------------------------------------------------------------------------------------------
TOL32 = 1e-5


@pytest.mark.parametrize('name', ['tiny', 'tiny_pool', 'tiny96', 'emnist', 'in64'])
def test_forward_fp32_class_matches_reference_golden(cuda_dev, name):
    """model.operand_format = FMT_FP32X: every tensor-core product as three fp16 MMAs (hi hi + lo hi + hi lo), accurate
    SiLU / softmax exponentials, fp32 everywhere else. Bar: max|out - ref| / max|ref| <= 1e-5 against the reference's own fp32
    output (the reference differs from an fp64 evaluation of itself by 5e-7..9e-7)."""
    from nice_diffusion_b200 import _lib as L
    cfg, sd, model = build(name, 1234, cuda_dev)
    model.operand_format = L.FMT_FP32X
    g = np.load(os.path.join(GOLD, 'unet_{}.npz'.format(name)))
    x = torch.from_numpy(g['x']).to(cuda_dev)
    t = torch.from_numpy(g['t']).to(cuda_dev)
    y = torch.from_numpy(g['y']).to(cuda_dev) if 'y' in g.files else None
    with torch.no_grad():
        out = model(x, t, y) if y is not None else model(x, t)
    ref = torch.from_numpy(g['out'])
    err = norm_err(out.cpu(), ref)
    print(name, 'fp32-class max|d|/max|ref| =', err)
    assert err < TOL32
    h = ref.shape[1] // 2
    assert norm_err(out.cpu()[:, :h], ref[:, :h]) < TOL32 and norm_err(out.cpu()[:, h:], ref[:, h:]) < TOL32
    # switching back restores the 16-bit plan (and its error level)
    model.operand_format = L.FMT_FP16
    with torch.no_grad():
        out16 = model(x, t, y) if y is not None else model(x, t)
    e16 = norm_err(out16.cpu(), ref)
    assert TOL32 < e16 < TOL
