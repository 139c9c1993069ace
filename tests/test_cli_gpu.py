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
