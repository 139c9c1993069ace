"""Preset model / diffusion argument dicts, same names and values as the reference's
nicediffusion/default_args.py (EMNIST :4-10, 64x64 :15-21, 128x128 :26-32, 256x256 :37-43)."""


def _diffusion(schedule, ddim, guidance):
    return {'rescaled_num_steps': 25, 'original_num_steps': 1000, 'use_ddim': ddim, 'ddim_eta': 0.0,
            'beta_schedule': schedule, 'sampling_var_type': 'learned_interpolation', 'classifier': None,
            'guidance_method': guidance, 'guidance_strength': 0.8, 'loss_type': 'hybrid'}


def _model(resolution, mult, width, blocks, in_ch, classes, attn=(8, 16, 32), heads=None, head_channels=None):
    args = {'resolution': resolution, 'attention_resolutions': attn, 'channel_mult': mult}
    if heads is not None:
        args['num_heads'] = heads
    if head_channels is not None:
        args['num_head_channels'] = head_channels
    args.update({'in_channels': in_ch, 'out_channels': 2 * in_ch, 'model_channels': width,
                 'num_res_blocks': blocks, 'split_qkv_first': True, 'dropout': 0.05,
                 'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': classes})
    return args


EMNIST_DIFFUSION_ARGS = _diffusion('cosine', False, 'classifier_free')
EMNIST_MODEL_ARGS = _model(28, (1, 2, 4), 64, 2, 1, 27, attn=(7, 14), heads=4)

OPENAI_64_DIFFUSION_ARGS = _diffusion('cosine', True, None)
OPENAI_64_MODEL_ARGS = _model(64, (1, 2, 3, 4), 192, 3, 3, 1000, head_channels=64)

OPENAI_128_DIFFUSION_ARGS = _diffusion('linear', True, None)
OPENAI_128_MODEL_ARGS = _model(128, (1, 1, 2, 3, 4), 256, 2, 3, 1000, heads=4)

OPENAI_256_DIFFUSION_ARGS = _diffusion('linear', True, None)
OPENAI_256_MODEL_ARGS = _model(256, (1, 1, 2, 2, 4, 4), 256, 2, 3, 1000, head_channels=64)
