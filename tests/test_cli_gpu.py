"""scripts/sample.py end to end on the GPU: the reference's CLI semantics (scripts/sample.py:16-184 of the reference) with a
small custom architecture and seeded random weights -- file naming, number of images, determinism of the Philox path, the
--start_img / --steps_to_do partial-denoise path, the refused --cpu flag."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUSTOM = ['-c', '--resolution', '16', '--model_channels', '64', '--channel_mult', '1/2', '--num_res_blocks', '1',
          '--attention_resolutions', '8', '--num_classes', '5', '--in_channels', '3', '--num_head_channels', '32',
          '--resblock_updown', '--use_adaptive_gn', '--rescaled_num_steps', '4', '--beta_schedule', 'cosine',
          '--sampling_var_type', 'learned_interpolation']


def run_cli(extra, cwd):
    cmd = [sys.executable, os.path.join(ROOT, 'scripts', 'sample.py'), '--model_path', 'custom_random.pt', '--random_weights',
           '--batch_size', '3', '--num_samples', '2'] + CUSTOM + extra
    return subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600)


def test_sample_cli_files_determinism_and_start_img(cuda_dev, tmp_path):
    import cv2
    a, b, c = tmp_path / 'a', tmp_path / 'b', tmp_path / 'c'
    for d in (a, b, c):
        d.mkdir()
    r = run_cli(['--noise', 'philox', '--seed', '7', '--labels', '1/3', '--save_path', str(a) + os.sep], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    names = sorted(os.path.basename(p) for p in glob.glob(str(a / '*.jpg')))
    # one label per batch, {label}_sample{k}.jpg, k counts the images of that label
    assert names == ['1_sample0.jpg', '1_sample1.jpg', '1_sample2.jpg', '3_sample0.jpg', '3_sample1.jpg', '3_sample2.jpg']
    img = cv2.imread(str(a / '1_sample0.jpg'))
    assert img.shape == (16, 16, 3) and img.std() > 0
    # same seed -> byte-identical files
    r = run_cli(['--noise', 'philox', '--seed', '7', '--labels', '1/3', '--save_path', str(b) + os.sep], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    for n in names:
        assert (a / n).read_bytes() == (b / n).read_bytes()
    # img2img entry: q-sample the start image to step steps_to_do * R // T and denoise from there
    start = tmp_path / 'start.png'
    cv2.imwrite(str(start), (np.random.RandomState(0).rand(20, 24, 3) * 255).astype(np.uint8))
    r = run_cli(['--seed', '7', '--labels', '2/2', '--start_img', str(start), '--steps_to_do', '500', '--save_path', str(c) + os.sep],
                str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(glob.glob(str(c / '2_sample*.jpg'))) == 6
    # no CPU implementation behind this CLI
    r = run_cli(['--cpu', '--save_path', str(c) + os.sep], str(tmp_path))
    assert r.returncode != 0 and 'no CPU implementation' in (r.stderr + r.stdout)


def test_sample_cli_grayscale_bytes_follow_the_reference_formula(cuda_dev, tmp_path):
    """One-channel models (EMNIST, BASELINE config 1): the reference inverts the float image, truncates to uint8 and inverts
    again (scripts/sample.py:98-100,165-171 of the reference), so the stored byte is 255 - trunc(255 - v). The CLI's arrays
    (one ndb_to_uint8 launch) must equal that formula evaluated with the reference's torch ops on the host, and the file
    must hold exactly that array pushed through the same writer."""
    import torch
    from nice_diffusion_b200.utils import images_to_uint8, reference_image_formula, save_image
    out_dir = tmp_path / 'g'
    out_dir.mkdir()
    raw = tmp_path / 'raw.npy'
    gray = [a if a != '3' or CUSTOM[i - 1] != '--in_channels' else '1' for i, a in enumerate(CUSTOM)]
    cmd = [sys.executable, os.path.join(ROOT, 'scripts', 'sample.py'), '--model_path', 'custom_random.pt', '--random_weights',
           '--batch_size', '3', '--num_samples', '1', '--noise', 'philox', '--seed', '11', '--labels', '2',
           '--save_path', str(out_dir) + os.sep, '--dump_raw', str(raw)] + gray
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    x = torch.from_numpy(np.load(raw))
    assert x.shape == (3, 1, 16, 16)
    want = reference_image_formula(x, 1)                               # host, the reference's own ops
    got = images_to_uint8(x.to(cuda_dev), 1)                           # the kernel the CLI used
    assert want.shape == (3, 16, 16) and want.dtype == np.uint8 and np.array_equal(got, want)
    assert (want != (((x + 1) * 127.5).clamp(0, 255)).to(torch.uint8)[:, 0].numpy()).any()   # differs from a single truncation
    for b in range(3):
        ref_file = tmp_path / 'ref{}.jpg'.format(b)
        save_image(str(ref_file), want[b])
        assert (out_dir / '2_sample{}.jpg'.format(b)).read_bytes() == ref_file.read_bytes()
    # colour path: same comparison on random data, edge values included
    v = torch.rand(2, 3, 8, 8) * 2.4 - 1.2
    v[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999999])
    assert np.array_equal(images_to_uint8(v.to(cuda_dev), 3), reference_image_formula(v, 3))
    g1 = torch.rand(4, 1, 8, 8) * 2.4 - 1.2
    assert np.array_equal(images_to_uint8(g1.to(cuda_dev), 1), reference_image_formula(g1, 1))
