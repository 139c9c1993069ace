"""`Trainer` -- drop-in for nicediffusion.trainer.Trainer (reference trainer.py:20-141): same constructor, `train()`,
`sample()`, `save()`, checkpoint tuple and file names -- plus what the reference lists as a TODO (trainer.py:8):
data-parallel training with one process per GPU.

Data parallelism: every rank holds a full replica and its own micro-batch; gradients are
averaged with NCCL all-reduce between `backward()` and `opt.step()`. `GradBuckets` lays the gradients out ONCE in a few
large flat buffers (`param.grad` are views into them), counts gradient arrivals with post-accumulate hooks during
backward and launches the all-reduce of a bucket asynchronously the moment its last gradient has landed, so the exchange
of the late layers' gradients overlaps the backward pass of the early ones. Buckets are sized for launch latency and
overlap, not for links: an NVSwitch box gives every GPU full bandwidth to every peer.

Behaviour of the reference that is reproduced on purpose (each is visible in tests/test_trainer*.py):
  * `Trainer.__init__` mutates the caller's `diffusion_args` to build a 250-step DDPM sampler (trainer.py:35);
  * without a checkpoint the EMA dictionary ALIASES the live parameters (trainer.py:55: `param.data`, no clone), so the
    "EMA" equals the weights; pass `ema_clone=True` (an addition) for a real exponential moving average;
  * with gradient accumulation N only every N-th micro-batch is back-propagated (trainer.py:93-96);
  * t is drawn in [0, original_num_steps) (trainer.py:88): the loss indexes the respaced tables with it, so training
    needs rescaled_num_steps == original_num_steps (otherwise IndexError, like the reference's gather);
  * EMNIST batches are transposed (trainer.py:79) and labels become the null class with probability 2 % under
    classifier-free guidance (trainer.py:84-85; the comment there says 1 %).
"""
import os

import numpy as np
import torch
import torch.optim as optim

from .diffusion import Diffusion


class GradBuckets:
    """Flat gradient buckets + overlapped all-reduce for a replicated model.

    Parameters are assigned to buckets in REVERSE registration order (the order backward produces their gradients);
    each bucket is one contiguous fp32 buffer and every `param.grad` is a view into it, so nothing is copied before or
    after the collective. `finish()` waits for the outstanding collectives and turns the sums into means."""

    def __init__(self, params, bucket_bytes=128 << 20, process_group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # dict(flat, params, pending, work)
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nbytes
        if cur:
            self._close(cur)
        self._handles = []
        for bi, b in enumerate(self.buckets):
            for p in b['params']:
                self._handles.append(p.register_post_accumulate_grad_hook(lambda _p, bi=bi: self._arrived(bi)))
        self.bytes = sum(b['flat'].numel() * 4 for b in self.buckets)

    def _close(self, plist):
        dev = plist[0].device
        flat = torch.zeros(sum(p.numel() for p in plist), dtype=torch.float32, device=dev)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)       # autograd accumulates into this view from now on
            off += p.numel()
        self.buckets.append(dict(flat=flat, params=plist, pending=len(plist), work=None))

    def _arrived(self, bi):
        b = self.buckets[bi]
        b['pending'] -= 1
        if b['pending'] == 0 and self.world > 1:
            # NCCL runs the collective on its own stream; it starts when the producing kernels of this bucket are done
            b['work'] = self.dist.all_reduce(b['flat'], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """After backward(): wait for the exchanges, average. Gradients that never arrived (unused parameters) still
        take part with their zeros so that every rank runs the same collectives."""
        for b in self.buckets:
            if self.world > 1:
                if b['work'] is None:
                    b['work'] = self.dist.all_reduce(b['flat'], op=self.dist.ReduceOp.SUM, group=self.group, async_op=True)
                b['work'].wait()
                b['flat'].div_(self.world)
            b['work'] = None
            b['pending'] = len(b['params'])

    def zero(self):
        for b in self.buckets:
            b['flat'].zero_()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


class Trainer:
    def __init__(self, model, diffusion_args, dataloader, iterations,
                 batch_size, lr, weight_decay, ema_rate=0.9999,
                 grad_accumulation=1,
                 checkpoint=(None, None, None, None),
                 print_every=None, sample_every=None, save_every=None, device=None,
                 ema_clone=False, bucket_bytes=128 << 20, checkpoint_dir='checkpoints'):
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu') if device is None else device
        self.model = model
        self.model.to(self.device)
        self.grad_accumulation = grad_accumulation

        self.train_diffusion = Diffusion(**diffusion_args, model=model, device=self.device)
        diffusion_args.update({'rescaled_num_steps': 250, 'use_ddim': False})      # the caller's dict, as the reference does
        self.sampling_diffusion = Diffusion(**diffusion_args, model=model, device=self.device)

        self.opt = optim.AdamW(self.model.parameters(), lr=lr, weight_decay=weight_decay, betas=(0.9, 0.999))
        self.loader = dataloader
        self.batch_size = batch_size
        self.checkpoint_dir = checkpoint_dir

        if any(c is not None for c in checkpoint):
            assert None not in checkpoint, \
                'please provide model, ema, and optimizer checkpoint paths and number of steps to resume with'
            self.model.load_state_dict(torch.load(checkpoint[0], map_location=self.device), strict=True)
            ema_state_dict = torch.load(checkpoint[1], map_location=self.device)
            self.opt.load_state_dict(torch.load(checkpoint[2], map_location=self.device))
            self.start_step = checkpoint[3]
            self.ema_params = {name: ema_state_dict[name] for name, _ in model.named_parameters()}
        else:
            self.start_step = 0
            self.ema_params = {name: (param.data.clone() if ema_clone else param.data)
                               for name, param in model.named_parameters()}
        for _, p in self.ema_params.items():
            p.requires_grad = False

        self.ema_rate = ema_rate
        self.iterations = iterations
        self.print_every = print_every
        self.sample_every = sample_every
        self.save_every = save_every

        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.buckets = GradBuckets(self.model.parameters(), bucket_bytes=bucket_bytes)
        self.last_loss = None

    # ------------------------------------------------------------------------------------------
    def train_step(self, batch, labels, step=0, t=None, noise=None):
        """One iteration of the reference loop body (trainer.py:69-101) on a batch already on the device. `t` / `noise`
        may be injected (tests); returns the scaled loss of this micro-batch."""
        self.model.train()
        if labels is not None:
            if self.train_diffusion.guidance == 'classifier_free' and np.random.randint(100) < 2:
                labels = torch.full_like(labels, fill_value=0).to(self.device)
            kwargs = {'y': labels}
        else:
            kwargs = None
        if t is None:
            t = torch.randint(low=0, high=self.train_diffusion.original_num_steps, size=(self.batch_size,),
                              device=self.device)
        loss = self.train_diffusion.loss(x_0=batch, t=t, kwargs=kwargs, noise=noise)
        loss = loss.mean() / self.grad_accumulation
        if (step + 1) % self.grad_accumulation == 0:
            loss.backward()                      # bucket all-reduces start inside, as gradients land
            self.buckets.finish()
            self.opt.step()
            self.buckets.zero()                  # == opt.zero_grad(set_to_none=False): the views stay in place
        # EMA: ema <- rate * ema + (1 - rate) * current (trainer.py:100-101)
        for name, param in self.model.named_parameters():
            self.ema_params[name].detach().mul_(self.ema_rate).add_(param.data, alpha=(1 - self.ema_rate))
        self.last_loss = loss.detach()
        return self.last_loss

    def train(self):
        running_loss = torch.tensor([0], dtype=torch.float, device=self.device)
        for step in range(self.iterations):
            batch, labels = next(self.loader)
            batch = batch.permute(0, 1, 3, 2).to(self.device)  # EMNIST is stored w,h (trainer.py:79)
            labels = labels.to(self.device) if labels is not None else None
            running_loss += self.train_step(batch, labels, step)
            if self.print_every is not None and step % self.print_every == 0 and self.rank == 0:
                print('Step #{}  ------------------------------------------\n\tLoss={}'.
                      format(step, running_loss.item() / self.print_every))
                running_loss = torch.tensor([0], dtype=torch.float, device=self.device)
            if self.sample_every is not None and step % self.sample_every == 0 and self.rank == 0:
                self.sample(4)
            if self.save_every is not None and step % self.save_every == 0:
                self.save(self.start_step + step)
        self.save(self.start_step + self.iterations)

    def sample(self, num_samples, show=True):
        """trainer.py:117-134: EMA weights, the 250-step DDPM sampler -- on the kernel plan (eval mode)."""
        self.model.eval()
        labels = torch.randint(low=0, high=self.model.num_classes, size=(num_samples,), device=self.device,
                               requires_grad=False)
        out = self.sampling_diffusion.denoise(kwargs={'y': labels}, batch_size=num_samples, ema_params=self.ema_params,
                                              progress=False)
        from .utils import images_to_uint8
        imgs = images_to_uint8(out, 3)           # trainer.py:121-122: no grey-scale inversion here, HWC uint8
        if show:
            from matplotlib import pyplot as plt
            for b in range(num_samples):
                plt.close('all')
                if self.model.in_channels == 1:
                    plt.imshow(imgs[b].astype(np.uint8), cmap='Greys')
                else:
                    plt.imshow(imgs[b].astype(np.uint8))
                plt.title('Label: {}'.format(labels[b].cpu().item()))
                plt.pause(0.0001)
                plt.waitforbuttonpress()
            plt.close('all')
        return imgs, labels

    def save(self, step):
        """trainer.py:136-141 (rank 0 writes; the replicas are identical)."""
        if self.rank != 0:
            return
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        torch.save(self.model.state_dict(), os.path.join(self.checkpoint_dir, '{}_model_params.pt'.format(step)))
        torch.save(self.opt.state_dict(), os.path.join(self.checkpoint_dir, '{}_opt_params.pt'.format(step)))
        torch.save(self.ema_params, os.path.join(self.checkpoint_dir, '{}_ema_params.pt'.format(step)))
        print('Saved checkpoint!')
