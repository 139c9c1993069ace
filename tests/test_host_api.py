"""CPU: host-side mirror of the reference API (no kernels run): state-dict schema, constructor defaults, argparse /
preset handling, schedule tables of the product `Diffusion`, C-ABI exports, no-CPU-fallback behaviour."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import torch

from oracle import CONFIGS
from oracle.unet import param_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def test_state_dict_schema_and_init():
    from nice_diffusion_b200.model import DiffusionModel
    for name, cfg in CONFIGS.items():
        if name in ('in128',):
            continue
        m = DiffusionModel(**cfg)
        sd = m.state_dict()
        want = param_shapes(cfg)   # pinned to the reference's key order by oracle/make_golden.py
        assert list(sd.keys()) == list(want.keys()), name
        assert all(tuple(sd[k].shape) == want[k] for k in want), name
        # zero-initialised tensors of the reference (model.py:177,253,448)
        for k, v in sd.items():
            if re.search(r'(out_conv|proj_out)\.(weight|bias)$', k) or k.startswith('out.2.'):
                assert float(v.abs().max()) == 0.0, k
        assert m.conditional == (cfg.get('num_classes') is not None)
        assert m.resolution == cfg['resolution'] and m.in_channels == cfg['in_channels']
        assert m.model_channels == cfg['model_channels'] and m.num_classes == cfg.get('num_classes')


def test_constructor_signature_matches_reference():
    from nice_diffusion_b200.model import DiffusionModel
    from nice_diffusion_b200.diffusion import Diffusion
    p = inspect.signature(DiffusionModel.__init__).parameters
    assert list(p)[1:] == ['resolution', 'in_channels', 'model_channels', 'out_channels', 'num_res_blocks',
                           'attention_resolutions', 'dropout', 'channel_mult', 'conv_resample', 'num_classes',
                           'num_heads', 'num_head_channels', 'resblock_updown', 'use_adaptive_gn', 'split_qkv_first',
                           'use_grad_checkpoints']
    assert p['dropout'].default == 0 and p['channel_mult'].default == (1, 2, 4, 8) and p['conv_resample'].default is True
    assert p['num_heads'].default == 1 and p['split_qkv_first'].default is True and p['resblock_updown'].default is False
    q = inspect.signature(Diffusion.__init__).parameters
    assert list(q)[1:] == ['model', 'original_num_steps', 'rescaled_num_steps', 'sampling_var_type', 'loss_type', 'betas',
                           'beta_schedule', 'guidance_method', 'guidance_strength', 'classifier', 'use_ddim', 'ddim_eta',
                           'device']
    d = inspect.signature(Diffusion.denoise).parameters
    assert list(d)[1:8] == ['x', 'kwargs', 'start_step', 'steps_to_do', 'batch_size', 'ema_params', 'progress']


def _tiny_model():
    from nice_diffusion_b200.model import DiffusionModel
    return DiffusionModel(**CONFIGS['tiny'])


@pytest.mark.parametrize('T,R,name', [(1000, 250, 'cosine'), (1000, 25, 'cosine'), (1000, 50, 'linear'),
                                      (1000, 300, 'cosine'), (100, 10, 'constant'), (4000, 100, 'cosine')])
def test_product_schedule_is_bit_exact(T, R, name):
    from nice_diffusion_b200.diffusion import Diffusion
    g = np.load(os.path.join(GOLD, 'schedule.npz'))
    d = Diffusion(model=_tiny_model(), original_num_steps=T, rescaled_num_steps=R, beta_schedule=name,
                  sampling_var_type='learned_interpolation', loss_type='hybrid', device=torch.device('cpu'))
    key = 'T{}_R{}_{}'.format(T, R, name)
    assert d.timestep_map.dtype == torch.long and np.array_equal(g[key + '/timestep_map'], d.timestep_map.numpy())
    for n in ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod', 'sqrt_one_minus_alphas_cumprod',
              'sqrt_reciprocal_alphas_cumprod', 'sqrt_reciprocal_alphas_minus_one_cumprod', 'posterior_mean_coef_x0',
              'posterior_mean_coef_xt', 'posterior_variance', 'log_posterior_var_clipped'):
        a = getattr(d, n)
        assert a.dtype == np.float64 and np.array_equal(g[key + '/' + n], a), n


def test_diffusion_argument_checks():
    from nice_diffusion_b200.diffusion import Diffusion
    m = _tiny_model()
    kw = dict(original_num_steps=1000, rescaled_num_steps=10, loss_type='hybrid', device=torch.device('cpu'))
    with pytest.raises(NotImplementedError):
        Diffusion(model=m, sampling_var_type='learned', guidance_method='bogus', **kw)
    with pytest.raises(NotImplementedError):
        Diffusion(model=m, sampling_var_type='nope', **kw)
    with pytest.raises(NotImplementedError):
        Diffusion(model=m, sampling_var_type='learned', **dict(kw, loss_type='nope'))
    with pytest.raises(AssertionError):
        Diffusion(model=m, sampling_var_type='learned', use_ddim=True, **kw)          # eta missing
    with pytest.raises(AssertionError):
        Diffusion(model=m, sampling_var_type='learned', betas=[0.1] * 5, **kw)        # wrong length
    from nice_diffusion_b200.model import DiffusionModel
    uncond = DiffusionModel(**CONFIGS['tiny_legacy'])
    with pytest.raises(AssertionError):
        Diffusion(model=uncond, sampling_var_type='small', guidance_method='classifier_free', **kw)
    d = Diffusion(model=m, sampling_var_type='learned', **kw)
    assert not m.training
    # the training loss is a CUDA kernel: on a CPU tensor it fails loudly (no CPU fallback), and a step index outside the
    # respaced schedule raises like the reference's gather
    from nice_diffusion_b200 import _lib
    with pytest.raises(_lib.NdbError):
        d.loss(torch.zeros(1, 3, 16, 16), torch.zeros(1))
    with pytest.raises(IndexError):
        d.loss(torch.zeros(1, 3, 16, 16), torch.full((1,), 10.0))


def test_no_cpu_fallback():
    from nice_diffusion_b200 import _lib
    m = _tiny_model().eval()
    with pytest.raises(_lib.NdbError):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))        # y missing: same assert as the reference


def test_argparser_presets_and_custom():
    from nice_diffusion_b200.utils import make_argparser, get_dicts_from_args
    from nice_diffusion_b200 import default_args as da
    p = make_argparser('diff_sample')
    other, m, d = get_dicts_from_args(p.parse_args(
        '--model_path models/64x64_diffusion.pt --batch_size 4 --num_samples 2 --labels 3/7 --seed 1'.split()))
    assert m['model_channels'] == 192 and m['num_classes'] == 1000 and m['out_channels'] == 6
    assert d == dict(da.OPENAI_64_DIFFUSION_ARGS) and other['labels'] == [3, 7] and other['seed'] == 1
    # EMNIST preset keeps 27 classes although it uses CFG (early return, utils.py:192-196)
    other, m, d = get_dicts_from_args(p.parse_args('--model_path models/EMNIST_model_params.pt --batch_size 16 --num_samples 1'.split()))
    assert m['num_classes'] == 27 and d['guidance_method'] == 'classifier_free' and other['labels'] == ''
    with pytest.raises(NotImplementedError):
        get_dicts_from_args(p.parse_args('--model_path foo.pt --batch_size 1 --num_samples 1'.split()))
    # custom path: '/'-separated lists, out_channels doubling, CFG adds the null class
    args = p.parse_args(('--model_path x.pt -c --batch_size 1 --num_samples 1 --resolution 32 --model_channels 64 '
                         '--channel_mult 1/2/2 --num_res_blocks 2 --attention_resolutions 8/16 --num_classes 10 '
                         '--rescaled_num_steps 50 --beta_schedule cosine --sampling_var_type learned '
                         '--guidance_method classifier_free --guidance_strength 1.0').split())
    other, m, d = get_dicts_from_args(args)
    assert m['channel_mult'] == [1, 2, 2] and m['attention_resolutions'] == [8, 16]
    assert m['out_channels'] == 6 and m['num_classes'] == 11 and m['num_heads'] == 4
    args = p.parse_args(('--model_path x.pt -c --batch_size 1 --num_samples 1 --resolution 32 --model_channels 64 '
                         '--channel_mult 1 --num_res_blocks 1 --attention_resolutions 8 '
                         '--rescaled_num_steps 50 --beta_schedule linear --sampling_var_type small').split())
    assert get_dicts_from_args(args)[1]['out_channels'] == 3
    with pytest.raises(NotImplementedError):   # classifier path is refused like the reference (utils.py:171-172)
        get_dicts_from_args(p.parse_args(('--model_path 64x64.pt --batch_size 1 --num_samples 1 --num_classes 5 '
                                          '--guidance_method classifier --classifier_path c.pt').split()))
    t = make_argparser('diff_train')
    with pytest.raises(SystemExit):
        t.parse_args([])
    with pytest.raises(NotImplementedError):
        make_argparser('other')


def test_convert_state_dict():
    from nice_diffusion_b200.utils import convert_state_dict
    sd = {'input_blocks.1.0.in_layers.0.weight': 1, 'output_blocks.2.1.qkv.weight': 2, 'time_embed.0.bias': 3,
          'label_emb.weight': 4, 'middle_block.0.emb_layers.1.weight': 5, 'input_blocks.4.0.out_layers.3.bias': 6,
          'output_blocks.0.0.skip_connection.weight': 7, 'out.2.weight': 8, 'input_blocks.1.0.in_layers.2.weight': 9,
          'input_blocks.1.0.out_layers.0.weight': 10}
    out = convert_state_dict(dict(sd))
    assert list(out.keys()) == ['downsampling.1.0.in_norm.weight', 'upsampling.2.1.qkv_nin.weight', 'step_embed.0.bias',
                                'class_embedding.weight', 'middle_block.0.step_embedding.weight',
                                'downsampling.4.0.out_conv.bias', 'upsampling.0.0.skip.weight', 'out.2.weight',
                                'downsampling.1.0.in_conv.weight', 'downsampling.1.0.out_norm.weight']
    assert list(out.values()) == list(range(1, 11))


def test_c_abi_library_exports_every_declared_symbol():
    from nice_diffusion_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'ndb200.h')).read()
    declared = set(re.findall(r'\b(ndb_[a-z0-9_]+)\s*\(', header))
    declared -= {n for n in declared if re.search(r'typedef struct[^;]*\b' + n + r'\b', header)}
    assert declared, 'no declarations parsed'
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), 'libndb200.so does not export ' + name
        assert name in _lib.SIGNATURES, 'ctypes binding missing for ' + name
    assert set(_lib.SIGNATURES) == declared
    assert lib.ndb_version() == 1
    # struct layouts the Python side mirrors (sizes are part of the ABI)
    assert ctypes.sizeof(_lib.ConvSegment) == 24
    assert _lib.MAX_SEGMENTS == int(re.search(r'#define NDB_MAX_SEGMENTS (\d+)', header).group(1))
    assert _lib.TAB_COUNT == int(re.search(r'#define NDB_TAB_COUNT (\d+)', header).group(1))


def test_invalid_arguments_return_error_codes_without_a_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU box."""
    from nice_diffusion_b200 import _lib
    lib = _lib.load()
    p = _lib.ConvParams()
    assert lib.ndb_conv_gemm(ctypes.byref(p), None) == _lib.ERR_INVALID
    assert lib.ndb_conv_gemm(None, None) == _lib.ERR_INVALID
    assert lib.ndb_attention(None, None, 1, 64, 1, 64, 0, 0, None) == _lib.ERR_INVALID
    assert lib.ndb_gn_apply(None, None) == _lib.ERR_INVALID
    assert lib.ndb_sampler_step(None, None) == _lib.ERR_INVALID
    assert lib.ndb_gn_stats_workspace_bytes(2, 64, 64) > 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, the CLI or the C sources may reference it (bench.py may,
    for its cpu_baseline / --impl reference legs only)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pat = re.compile(r'^\s*(from|import)\s+oracle\b', re.M)
    checked = 0
    for base, _, files in os.walk(os.path.join(root, 'nice-diffusion_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(base, f), errors='ignore').read()
                assert not pat.search(text), os.path.join(base, f)
                checked += 1
    for f in ('scripts/sample.py', 'nice_diffusion_b200.py'):
        assert not pat.search(open(os.path.join(root, f)).read()), f
    assert checked > 10


def test_bench_batch_follows_the_named_configuration():
    """bench.py: ImageNet-64 (BASELINE config 2: "batch 256 sharded over 8 x B200") runs the named batch on the GPUs of the
    run -- 256 / N per GPU, strong scaling; --scaling weak and an explicit --batch keep the per-GPU batch fixed."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('ndb_bench', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world, b in ((1, 256), (2, 128), (4, 64), (8, 32)):
        assert bench.per_gpu_batch('in64', world) == (b, 'strong')
        assert bench.per_gpu_batch('in64', world, None, 'weak') == (32, 'weak')
        assert bench.per_gpu_batch('in64', world, 48) == (48, 'weak')
        assert bench.per_gpu_batch('emnist', world) == (16, 'weak')
        assert bench.per_gpu_batch('in128', world) == (8, 'weak')
        assert bench.per_gpu_batch('in64_clf', world) == (32, 'weak')
    with pytest.raises(SystemExit):
        bench.per_gpu_batch('in64', 3)


def test_bench_traffic_capture_matches_the_batch_of_the_run(tmp_path):
    """`roofline.traffic` comes from a committed ncu capture: only one taken at the run's per-GPU batch may be used."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('ndb_bench2', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    (tmp_path / 'r01_conv_traffic.json').write_text(json.dumps({'launches': 118, 'dram_bytes_per_launch_avg': 1.0}))
    (tmp_path / 'r02_conv_traffic.json').write_text(json.dumps({'launches': 118, 'dram_bytes_per_launch_avg': 2.0}))
    (tmp_path / 'r02b256_conv_traffic.json').write_text(
        json.dumps({'launches': 118, 'dram_bytes_per_launch_avg': 9.0, 'per_gpu_batch': 256}))
    assert bench.conv_traffic_capture('in64', 32, str(tmp_path))['dram_bytes_per_launch_avg'] == 2.0     # newest tag wins
    assert bench.conv_traffic_capture('in64', 256, str(tmp_path))['file'] == 'r02b256_conv_traffic.json'
    assert bench.conv_traffic_capture('in64', 64, str(tmp_path)) is None
    assert bench.conv_traffic_capture('emnist', 32, str(tmp_path)) is None
    # the committed captures of this tree
    assert bench.conv_traffic_capture('in64', 256)['per_gpu_batch'] == 256
    assert bench.conv_traffic_capture('in64', 32) is not None
