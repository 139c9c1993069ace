"""Per-launch table of one denoising step (eager launches, CUDA events around each): desc, time, TFLOP/s or GB/s."""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nice_diffusion_b200 import ops
from nice_diffusion_b200.diffusion import Diffusion
from nice_diffusion_b200.model import DiffusionModel
from nice_diffusion_b200.utils import randomize_zero_init_

ap = argparse.ArgumentParser(); ap.add_argument('--config', default='in64'); ap.add_argument('--batch', type=int, default=None)
a = ap.parse_args()
cname, dkw, B, per = bench.WORKLOADS[a.config]
B = a.batch or B
dev = torch.device('cuda:0')
model = randomize_zero_init_(DiffusionModel(**bench.model_config(a.config))).to(dev).eval()
diff = Diffusion(model=model, device=dev, **dkw); diff.noise = 'philox'
x = torch.randn(B, model.in_channels, model.resolution, model.resolution, device=dev)
y = torch.randint(1, 20, (B,), device=dev)
runner = diff.make_step_runner(x, {'y': y})
runner.eng.serial = True     # one stream: per-launch events must not overlap
runner._launch(); runner._launch(); torch.cuda.synchronize()
reps = 3
acc = collections.OrderedDict()
for _ in range(reps):
    ops.PROFILE = []
    runner._launch(); torch.cuda.synchronize()
    for i, (name, e0, e1, flops, desc) in enumerate(ops.PROFILE):
        k = (name, desc)
        d = acc.setdefault(k, [0, 0.0, 0])
        d[0] += 1; d[1] += e0.elapsed_time(e1); d[2] += flops
    ops.PROFILE = None
tot = sum(d[1] for d in acc.values()) / reps
print('total per step (sum of kernel times): %.3f ms' % tot)
byname = collections.Counter()
for (name, desc), d in acc.items():
    byname[name] += d[1] / reps
print({k: round(v, 3) for k, v in byname.items()})
print('%-12s %-42s %5s %9s %9s %8s' % ('kernel', 'shape', 'n', 'ms/step', 'us/launch', 'TFLOP/s'))
for (name, desc), d in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    n = d[0] // reps
    ms = d[1] / reps
    tf = (d[2] / reps) / (ms * 1e-3) / 1e12 if d[2] else 0
    print('%-12s %-42s %5d %9.3f %9.1f %8.1f' % (name, desc, n, ms, 1e3 * ms / n, tf))
