"""Experiment: one B=32 step graph vs two concurrent B=16 step graphs on two streams (do small layers overlap?)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nice_diffusion_b200.diffusion import Diffusion
from nice_diffusion_b200.model import DiffusionModel
from nice_diffusion_b200.utils import randomize_zero_init_

cname, dkw, B, per = bench.WORKLOADS['in64']
dev = torch.device('cuda:0')
model = randomize_zero_init_(DiffusionModel(**bench.model_config('in64'))).to(dev).eval()

def make(model, b, seed):
    diff = Diffusion(model=model, device=dev, **dkw); diff.noise = 'philox'; diff.seed = seed
    x = torch.randn(b, 3, 64, 64, device=dev)
    y = torch.randint(1, 20, (b,), device=dev)
    return diff.make_step_runner(x, {'y': y})

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

r32 = make(model, 32, 0)
print('B=32 single graph: %.3f ms/step' % timeit(r32.step))
for lanes in (2, 4):
    b = 32 // lanes
    models = [model] + [copy.deepcopy(model) for _ in range(lanes - 1)]
    runners = [make(m, b, i) for i, m in enumerate(models)]
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    for r in runners: r.step()
    torch.cuda.synchronize()
    def step_all():
        cur = torch.cuda.current_stream()
        for s, r in zip(streams, runners):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                r.step()
        for s in streams:
            cur.wait_stream(s)
    print('%d lanes of B=%d on %d streams: %.3f ms/step' % (lanes, b, lanes, timeit(step_all)))
    r1 = runners[0]
    print('   (one B=%d graph alone: %.3f ms/step)' % (b, timeit(r1.step)))
