"""ORACLE -- CPU restatement of the reference's sampling hot path. TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``nice-diffusion_b200/``) imports this. It may be used by
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs, and there only as the checker / the CPU arm -- never as the thing shipped.

Pinning: the reference (edogariu/nice-diffusion) has no tests, fixtures or golden vectors
(SURVEY.md section 4). The oracle is therefore pinned against the reference *itself*, imported from
/root/reference in the build container by ``oracle/make_golden.py`` (bit-exact on CPU fp32 for the
UNet forward, the schedule tables and fixed-noise trajectories), and the outputs are committed
as fixtures under ``tests/golden/`` so the check travels to the GPU box where /root/reference
does not exist.
"""

CONFIGS = {
    # reference presets (nicediffusion/default_args.py:7-10,18-21,29-32)
    'emnist': {'resolution': 28, 'attention_resolutions': (7, 14), 'channel_mult': (1, 2, 4),
               'num_heads': 4, 'in_channels': 1, 'out_channels': 2, 'model_channels': 64,
               'num_res_blocks': 2, 'split_qkv_first': True, 'dropout': 0.05,
               'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': 27},
    'in64': {'resolution': 64, 'attention_resolutions': (8, 16, 32), 'channel_mult': (1, 2, 3, 4),
             'num_head_channels': 64, 'in_channels': 3, 'out_channels': 6, 'model_channels': 192,
             'num_res_blocks': 3, 'split_qkv_first': True, 'dropout': 0.05,
             'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': 1000},
    'in128': {'resolution': 128, 'attention_resolutions': (8, 16, 32), 'channel_mult': (1, 1, 2, 3, 4),
              'num_heads': 4, 'in_channels': 3, 'out_channels': 6, 'model_channels': 256,
              'num_res_blocks': 2, 'split_qkv_first': True, 'dropout': 0.05,
              'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': 1000},
    # small synthetic configs used by the fixtures (finish in well under a second on CPU)
    'tiny': {'resolution': 16, 'attention_resolutions': (8,), 'channel_mult': (1, 2),
             'num_head_channels': 64, 'in_channels': 3, 'out_channels': 6, 'model_channels': 64,
             'num_res_blocks': 1, 'split_qkv_first': True, 'dropout': 0.0,
             'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': 11},
    # exercises: legacy qkv order, plain (non-adaptive) GN, conv up/down-sampling, num_heads,
    # unconditional, out_channels == in_channels
    'tiny_legacy': {'resolution': 16, 'attention_resolutions': (8, 16), 'channel_mult': (1, 2),
                    'num_heads': 2, 'in_channels': 3, 'out_channels': 3, 'model_channels': 64,
                    'num_res_blocks': 1, 'split_qkv_first': False, 'dropout': 0.0,
                    'resblock_updown': False, 'conv_resample': True, 'use_adaptive_gn': False,
                    'num_classes': None},
    # widths that are multiples of 32 but not of 64 (the reference's --custom --model_channels 96): 96 / 192 channels, heads of 32
    'tiny96': {'resolution': 16, 'attention_resolutions': (8, 16), 'channel_mult': (1, 2),
               'num_head_channels': 32, 'in_channels': 3, 'out_channels': 6, 'model_channels': 96,
               'num_res_blocks': 1, 'split_qkv_first': True, 'dropout': 0.0,
               'resblock_updown': True, 'use_adaptive_gn': True, 'num_classes': 7},
    # avg-pool / nearest resampling without ResBlocks (conv_resample=False)
    'tiny_pool': {'resolution': 16, 'attention_resolutions': (), 'channel_mult': (1, 1),
                  'num_heads': 1, 'in_channels': 3, 'out_channels': 6, 'model_channels': 64,
                  'num_res_blocks': 1, 'split_qkv_first': True, 'dropout': 0.0,
                  'resblock_updown': False, 'conv_resample': False, 'use_adaptive_gn': True,
                  'num_classes': 5},
}
