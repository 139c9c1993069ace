"""Host-side configuration helpers with the reference's names and behaviour (nicediffusion/utils.py):
`make_argparser`, `get_dicts_from_args`, `convert_state_dict`, plus small helpers used by the sampler CLI.
No matplotlib dependency (the reference imports it at module top, utils.py:3)."""
import argparse
from collections import OrderedDict

from .default_args import (EMNIST_DIFFUSION_ARGS, EMNIST_MODEL_ARGS, OPENAI_64_DIFFUSION_ARGS, OPENAI_64_MODEL_ARGS,
                           OPENAI_128_DIFFUSION_ARGS, OPENAI_128_MODEL_ARGS, OPENAI_256_DIFFUSION_ARGS,
                           OPENAI_256_MODEL_ARGS)

_O, _R = '(optional)', '(required)'

# (flags, kwargs) tables -- same flag names, types and defaults as utils.py:28-142 of the reference
_SAMPLING = [
    (('--model_path',), dict(type=str, required=True, metavar=_R, help='relative file path of model state dict')),
    (('-c', '--custom'), dict(action='store_true', default=False,
                              help='whether to use a custom model/diffusion configuration')),
    (('--batch_size',), dict(type=int, required=True, metavar=_R, help='number of images per batch')),
    (('--num_samples',), dict(type=int, required=True, metavar=_R,
                              help='number of batches to sample. total images = num_samples * batch_size')),
    (('--upsample',), dict(action='store_true', default=False, help='4x Real-ESRGAN super-resolution (needs basicsr)')),
    (('--wordy', '-w'), dict(dest='wordy', action='store_true', default=False, help='print status')),
    (('--save_path',), dict(type=str, metavar=_O, default=None, help='where to save images (else display)')),
    (('--labels',), dict(type=str, metavar=_O, default='', help="'/'-separated labels, one per batch")),
    (('--start_img',), dict(type=str, metavar=_O, default=None, help='image to start denoising from')),
    (('--steps_to_do',), dict(type=int, metavar=_O, default=None,
                              help='steps of noise to apply to start_img and then remove')),
    (('--seed',), dict(type=int, metavar=_O, default=None, help='rng seed')),
    (('--cpu',), dict(action='store_true', default=False,
                      help='accepted for CLI compatibility; the ndb200 path has no CPU implementation and refuses it')),
]
_TRAINING = [
    (('--batch_size',), dict(type=int, required=True, metavar=_R)),
    (('--lr',), dict(type=float, required=True, metavar=_R)),
    (('--weight_decay',), dict(type=float, required=True, metavar=_R)),
    (('--iterations',), dict(type=int, required=True, metavar=_R)),
    (('--resume_step',), dict(type=int, metavar=_O, default=0)),
    (('--wordy', '-w'), dict(dest='wordy', action='store_true', default=False)),
    (('--save_every',), dict(type=int, metavar=_O, default=None)),
    (('--sample_every',), dict(type=int, metavar=_O, default=None)),
    (('--ema_rate',), dict(type=float, metavar=_O, default=0.9999)),
    (('--use_fp16',), dict(action='store_true', default=False)),
    (('--grad_accumulation',), dict(type=int, metavar=_O, default=1)),
    (('--seed',), dict(type=int, metavar=_O, default=None)),
]


def _model_table(req):
    m = _R if req else _O
    return [
        (('--resolution',), dict(type=int, required=req, metavar=m, default=None)),
        (('--model_channels',), dict(type=int, required=req, metavar=m, default=None)),
        (('--channel_mult',), dict(type=str, required=req, metavar=m, default=None)),
        (('--num_res_blocks',), dict(type=int, required=req, metavar=m, default=None)),
        (('--attention_resolutions',), dict(type=str, required=req, metavar=m, default=None)),
        (('--num_classes',), dict(type=int, default=None, metavar=_O)),
        (('--dropout',), dict(type=float, required=req, default=0.0, metavar=m)),
        (('--in_channels',), dict(type=int, default=3, metavar=_O)),
        (('--num_heads',), dict(type=int, default=4, metavar=_O)),
        (('--num_head_channels',), dict(type=int, default=None, metavar=_O)),
        (('--split_qkv_first',), dict(action='store_true', default=False)),
        (('--resblock_updown',), dict(action='store_true', default=False)),
        (('--use_adaptive_gn',), dict(action='store_true', default=False)),
    ]


def _diffusion_table(req):
    m = _R if req else _O
    return [
        (('--rescaled_num_steps',), dict(type=int, required=req, metavar=m, default=None)),
        (('--beta_schedule',), dict(type=str, required=req, metavar=m, default=None)),
        (('--sampling_var_type',), dict(type=str, required=req, metavar=m, default=None)),
        (('--use_ddim',), dict(action='store_true', default=False)),
        (('--ddim_eta',), dict(type=float, default=0.0, metavar=_O)),
        (('--original_num_steps',), dict(type=int, default=1000, metavar=_O)),
        (('--loss_type',), dict(type=str, required=req, default='hybrid', metavar=m)),
        (('--guidance_method',), dict(type=str, default=None, metavar=_O)),
        (('--guidance_strength',), dict(type=float, default=None, metavar=_O)),
        (('--classifier_path',), dict(type=str, default=None, metavar=_O)),
    ]


def make_argparser(prog):
    """utils.py:12-143: 'diff_sample' or 'diff_train'."""
    if prog == 'diff_sample':
        description, is_sample = 'Sample images generated from Diffusion Model.', True
    elif prog == 'diff_train':
        description, is_sample = 'Train Diffusion Model.', False
    else:
        raise NotImplementedError(prog)
    parser = argparse.ArgumentParser(prog=prog, description=description)
    groups = [('sampling arguments' if is_sample else 'training arguments', _SAMPLING if is_sample else _TRAINING),
              ('model arguments', _model_table(not is_sample)),
              ('diffusion arguments', _diffusion_table(not is_sample))]
    for title, table in groups:
        grp = parser.add_argument_group(title)
        for flags, kw in table:
            kw = dict(kw)
            if 'action' in kw:
                kw.pop('metavar', None)
            grp.add_argument(*flags, **kw)
    return parser


MODEL_KEYS = ('resolution', 'attention_resolutions', 'channel_mult', 'num_res_blocks', 'model_channels', 'num_heads',
              'num_head_channels', 'in_channels', 'out_channels', 'split_qkv_first', 'dropout', 'resblock_updown',
              'use_adaptive_gn', 'num_classes')
DIFFUSION_KEYS = ('rescaled_num_steps', 'original_num_steps', 'use_ddim', 'ddim_eta', 'beta_schedule',
                  'sampling_var_type', 'classifier', 'guidance_method', 'guidance_strength', 'loss_type')
_PRESETS = (('64x64', OPENAI_64_MODEL_ARGS, OPENAI_64_DIFFUSION_ARGS),
            ('128x128', OPENAI_128_MODEL_ARGS, OPENAI_128_DIFFUSION_ARGS),
            ('256x256', OPENAI_256_MODEL_ARGS, OPENAI_256_DIFFUSION_ARGS),
            ('EMNIST', EMNIST_MODEL_ARGS, EMNIST_DIFFUSION_ARGS))


def _split_ints(text):
    return [int(i) for i in text.split('/')]


def get_dicts_from_args(args):
    """utils.py:146-214 -> (other_args, model_args, diffusion_args), including its quirks: the preset branch
    returns early (no list parsing, no out_channels derivation, no CFG `num_classes += 1`)."""
    model_args, diff_args, other = {}, {}, {}
    for key, val in vars(args).items():
        (model_args if key in MODEL_KEYS else diff_args if key in DIFFUSION_KEYS else other)[key] = val

    assert diff_args['guidance_method'] is None or model_args['num_classes'] is not None, \
        'use guidance only for conditional models'
    assert (diff_args['guidance_method'] == 'classifier') == (other['classifier_path'] is not None)
    if other['classifier_path'] is not None:
        raise NotImplementedError('i have not yet implemented a noisy classifier, sorry')

    if 'custom' in other:  # sampling mode
        if other['custom']:
            needed = [model_args[k] for k in ('resolution', 'model_channels', 'channel_mult', 'num_res_blocks',
                                              'attention_resolutions')]
            needed += [diff_args[k] for k in ('rescaled_num_steps', 'sampling_var_type', 'beta_schedule')]
            if not all(needed):
                raise argparse.ArgumentError(None, 'if the model is custom, all the configuration things must be specified.')
        else:
            path = other['model_path']
            for tag, m, d in _PRESETS:
                if tag in path:
                    break
            else:
                raise NotImplementedError(path, 'this is not a default model')
            model_args.update(m)
            diff_args.update(d)
            if 'labels' in other and len(other['labels']) > 0:
                other['labels'] = _split_ints(other['labels'])
            return other, model_args, diff_args

    if 'labels' in other and len(other['labels']) > 0:
        other['labels'] = _split_ints(other['labels'])
    model_args['attention_resolutions'] = _split_ints(model_args['attention_resolutions'])
    model_args['channel_mult'] = _split_ints(model_args['channel_mult'])
    learned = diff_args['sampling_var_type'] in ('learned', 'learned_interpolation')
    model_args['out_channels'] = model_args['in_channels'] * (2 if learned else 1)
    if diff_args['guidance_method'] == 'classifier_free':
        model_args['num_classes'] += 1
    return other, model_args, diff_args


_RENAMES = (('input_blocks', 'downsampling'), ('output_blocks', 'upsampling'), ('in_layers.0', 'in_norm'),
            ('in_layers.2', 'in_conv'), ('emb_layers.1', 'step_embedding'), ('out_layers.0', 'out_norm'),
            ('out_layers.3', 'out_conv'), ('skip_connection', 'skip'), ('time_embed', 'step_embed'),
            ('qkv', 'qkv_nin'), ('label_emb', 'class_embedding'))


def convert_state_dict(sd):
    """openai/guided-diffusion key names -> this model's (utils.py:265-292). Returns a new OrderedDict."""
    out = OrderedDict()
    for key, val in sd.items():
        for old, new in _RENAMES:
            key = key.replace(old, new)
        out[key] = val
    return out


def grab_checkpoint(step):
    """utils.py:301-306."""
    return ('checkpoints/{}_model_params.pt'.format(step), 'checkpoints/{}_ema_params.pt'.format(step),
            'checkpoints/{}_opt_params.pt'.format(step), step)


def randomize_zero_init_(model, seed=0):
    """The reference zero-initialises every out_conv / proj_out / head conv (model.py:507-510), so a fresh model
    outputs zeros. Benchmarks and smoke tests on random-init weights re-draw those tensors (seeded) so that
    every kernel sees non-trivial data: weights N(0, 1/fan_in), biases N(0, 0.02^2)."""
    import torch
    g = torch.Generator(device='cpu').manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if float(p.detach().abs().max()) == 0.0:
                if p.dim() > 1:
                    fan_in = p[0].numel()
                    p.copy_((torch.randn(p.shape, generator=g) / fan_in ** 0.5).to(p.device))
                else:
                    p.copy_((0.02 * torch.randn(p.shape, generator=g)).to(p.device))
    return model


def images_to_uint8(out, in_channels):
    """Sampler output [-1,1] NCHW (device) -> the uint8 arrays the reference CLI passes to `plt.imsave`
    (scripts/sample.py:94-100,165-171): [B,H,W,3] for colour models; for one-channel models the reference inverts the
    float image, truncates and inverts again, which leaves [B,H,W] = 255 - trunc(255 - v). One ndb_to_uint8 launch."""
    from . import ops
    gray = in_channels == 1
    img8 = ops.to_uint8(out.float(), gray_invert=gray).cpu().numpy()
    return img8[..., 0] if gray else img8


def reference_image_formula(out_cpu, in_channels):
    """The same conversion spelled with the reference's own torch ops on a CPU tensor (what the tests compare
    `images_to_uint8` against, byte for byte)."""
    import torch
    out = ((out_cpu + 1) * 127.5).clamp(0, 255)
    if in_channels == 1:
        out = torch.stack((255 - out.squeeze(1),) * 3, dim=1)
    out = out.to(torch.uint8).permute(0, 2, 3, 1).detach().numpy()
    if in_channels == 1:
        out = 255 - out[..., 0]
    return out


def save_image(path, img8):
    """`plt.imsave(path, img)` like the reference (sample.py:180) when matplotlib is installed -- for [H,W] arrays that
    call applies matplotlib's default colour map, which is what the reference's EMNIST files contain; without
    matplotlib the same array is written with OpenCV (RGB -> BGR, grey as is)."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        plt = None
    if plt is not None:
        plt.imsave(path, img8)
        return
    import cv2
    import numpy as np
    cv2.imwrite(path, np.ascontiguousarray(img8 if img8.ndim == 2 else img8[..., ::-1]))
