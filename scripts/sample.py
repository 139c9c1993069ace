#!/usr/bin/env python
"""Sampling CLI -- same flags and behaviour as the reference's scripts/sample.py (argv re-split, presets chosen by
--model_path, --num_samples = number of BATCHES, one label per batch, [-1,1] -> [0,255], grayscale inversion, file
naming), running on the ndb200 kernels. Additions: runs under torchrun with one process per GPU, every batch sharded
over the ranks by global sample index (no collective; each rank writes its own files), `--noise philox` for sampling
that is invariant to the GPU count, `--random_weights` to run a preset architecture without a checkpoint.

    python scripts/sample.py --model_path models/64x64_diffusion.pt --batch_size 32 --num_samples 1 --save_path out/
    python -m torch.distributed.run --nproc-per-node 8 scripts/sample.py ... --batch_size 256 ...
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from nice_diffusion_b200.diffusion import Diffusion
from nice_diffusion_b200.model import DiffusionModel
from nice_diffusion_b200.sharding import global_indices, output_names, shard_range, world_info
from nice_diffusion_b200.utils import (get_dicts_from_args, images_to_uint8, make_argparser, randomize_zero_init_,
                                       save_image)


def main():
    # the reference re-splits argv on spaces (sample.py:18-21)
    sys.argv = [a for chunk in sys.argv for a in chunk.split(' ')]
    parser = make_argparser('diff_sample')
    parser.add_argument('--noise', choices=['torch', 'philox'], default=None,
                        help="per-step noise source: 'torch' = torch.randn_like like the reference (default on one GPU); "
                             "'philox' = in-kernel Philox keyed by global sample index (default under torchrun)")
    parser.add_argument('--random_weights', action='store_true', default=False,
                        help='use seeded random weights of the architecture instead of loading --model_path')
    parser.add_argument('--dump_raw', type=str, default=None,
                        help='also write the float [-1,1] sampler outputs of this rank to this .npy file (verification)')
    args = parser.parse_args()
    noise_mode, random_weights, dump_raw = args.noise, args.random_weights, args.dump_raw
    del args.noise, args.random_weights, args.dump_raw
    other, model_args, diff_args = get_dicts_from_args(args)

    rank, world, local_rank = world_info()
    if other['cpu'] or not torch.cuda.is_available():
        raise SystemExit('sample.py: the ndb200 path needs a B200 -- there is no CPU implementation '
                         '(use the reference for --cpu)')
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    seed = other['seed']
    if seed is not None:
        torch.manual_seed(seed)
    wordy = other['wordy'] and rank == 0
    n_batches, B = other['num_samples'], other['batch_size']
    labels_arg = other['labels']
    conditional = model_args['num_classes'] is not None
    save_path = other['save_path']
    if noise_mode is None:
        noise_mode = 'philox' if world > 1 else 'torch'
    if world > 1 and noise_mode != 'philox':
        raise SystemExit('sample.py: multi-GPU sampling needs --noise philox (noise keyed by global sample index)')

    model = DiffusionModel(**model_args)
    if random_weights:
        randomize_zero_init_(model, seed=0 if seed is None else seed)
    else:
        model.load_state_dict(torch.load(other['model_path'], map_location='cpu'), strict=True)
    model.to(device).eval()
    if wordy:
        print('Model made from {} with {} parameters! :)'.format(other['model_path'],
                                                                 sum(p.numel() for p in model.parameters())))
        print('Starting Diffusion! There are {} samples of {} images each'.format(n_batches, B))
    diffusion = Diffusion(model=model, **diff_args, device=device)
    diffusion.noise = noise_mode
    diffusion.seed = 0 if seed is None else seed

    start_img, steps_to_do = other['start_img'], other['steps_to_do']
    if start_img is not None and steps_to_do is not None:
        from cv2 import imread, resize
        img = imread(start_img)
        img = resize(img, dsize=(model_args['resolution'], model_args['resolution'])) / 127.5 - 1
        img = torch.from_numpy(img).permute(2, 0, 1)[[2, 1, 0], :, :].float()     # BGR -> RGB
        start_img = img[None].repeat(B, 1, 1, 1)
    if conditional and len(labels_arg) != 0:
        assert len(labels_arg) == n_batches, 'please provide NUM_SAMPLES={} labels'.format(n_batches)

    # host-side randomness is drawn for the FULL batch on every rank (same generator state everywhere) and sliced
    gen = torch.Generator(device='cpu')
    gen.manual_seed(torch.initial_seed() if seed is None else seed)
    results = []   # (global ids, uint8 HWC images, labels)
    raw = []
    all_labels = []
    lo, hi = shard_range(B, rank, world)
    with torch.no_grad():
        for i in range(n_batches):
            ids = torch.tensor(list(global_indices(i, B, rank, world)), dtype=torch.long, device=device)
            diffusion.sample_ids = ids
            if start_img is None or steps_to_do is None:
                steps = diff_args['rescaled_num_steps']
                if noise_mode == 'philox':
                    data = None                                   # x_T from Philox(seed, id, step = R)
                else:
                    data = torch.randn([B, model_args['in_channels'], model_args['resolution'],
                                        model_args['resolution']], generator=gen)[lo:hi].to(device)
            else:
                steps = steps_to_do * diff_args['rescaled_num_steps'] // diff_args['original_num_steps']
                nz = torch.randn(start_img.shape, generator=gen)[lo:hi].to(device)
                data = diffusion.diffuse(x_0=start_img[lo:hi].to(device), steps_to_do=steps, noise=nz)
            if conditional:
                if len(labels_arg) == 0:
                    full = torch.randint(low=0, high=model_args['num_classes'], size=(B,), generator=gen)
                else:
                    full = torch.full(size=(B,), fill_value=labels_arg[i])
                all_labels += full.tolist()
                labels = full[lo:hi].to(device)
            else:
                all_labels += [None] * B
                labels = None
            if wordy:
                print('Denoising sample {}! :)'.format(i + 1))
            out = diffusion.denoise(x=data, kwargs={'y': labels}, batch_size=hi - lo, progress=wordy, steps_to_do=steps)
            # [-1,1] -> the uint8 HWC arrays the reference hands to plt.imsave (sample.py:94,98-100,165-171), one kernel
            results.append((list(range(i * B + lo, i * B + hi)), images_to_uint8(out, model_args['in_channels'])))
            raw.append(out.cpu().numpy())

    if other['upsample']:
        raise SystemExit('--upsample needs basicsr + RealESRGAN weights, which are outside this package (SURVEY.md section 2 #10)')
    names = output_names(all_labels)
    if save_path is None:
        import matplotlib.pyplot as plt            # display path needs matplotlib like the reference
        for ids, imgs in results:
            for gid, im in zip(ids, imgs):
                plt.imshow(im, cmap='gray' if im.ndim == 2 else None)
                plt.title('Output Image' + ('' if all_labels[gid] is None else ', Label={}'.format(all_labels[gid])))
                plt.waitforbuttonpress()
    else:
        if wordy:
            print('Saving {} generated images to \'{}\'!'.format(n_batches * B, save_path))
        for ids, imgs in results:
            for gid, im in zip(ids, imgs):
                save_image(save_path + names[gid], im)          # SAVE_PATH + filename, as the reference
    if dump_raw is not None:
        np.save(dump_raw, np.concatenate(raw, 0))
    if wordy:
        print('Done! have a nice day')


if __name__ == '__main__':
    main()
