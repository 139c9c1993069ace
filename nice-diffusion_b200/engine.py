"""UNet execution plan: turns a `DiffusionModel` parameter tree into a static list of ndb200 kernel launches
for one (device, batch size), with all intermediates preallocated, weights packed for the tensor cores, and the
whole forward captured in a CUDA graph.

Dataflow per block (reference: model.py:188-211 ResidualBlock, :260-291 AttentionBlock):

  ResBlock   gn_apply[in_norm, SiLU, up/down, concat] -> conv3x3 GEMM (+bias; epilogue emits GroupNorm statistics)
             -> gn_apply[out_norm, AdaGN scale/shift, SiLU] -> conv3x3 GEMM (+bias, + residual x | + 1x1 skip as an
             extra K segment, x_resample folded into the epilogue; epilogue emits statistics for the next norm)
  Attention  gn_apply -> qkv GEMM (fp16 out) -> fused attention -> proj GEMM (+bias +x, statistics)

The residual stream is fp32 NHWC and every stream tensor travels with its per-channel (sum, sumsq) statistics, written
by the epilogue of the GEMM that produced it -- there is no separate statistics pass except for the stem output.
Every tensor-core operand is written once in 16-bit NHWC by the consuming norm's apply kernel. `torch.cat` of the up
path never materialises in fp32: the apply kernel reads both sources (model.py:474). All step-embedding Linear
layers run as ONE GEMM per forward.
"""
import contextlib
import gc
import os

import torch

from . import _lib as L
from . import ops
from .model import (AttentionBlock, Downsample, ResidualBlock, Upsample, timestep_embedding_freqs)


@contextlib.contextmanager
def capture_guard():
    """Around every CUDA-graph capture: Python's cyclic garbage collector must not run INSIDE a capture -- if it finalises
    an old torch.cuda.CUDAGraph there (a dropped plan, a previous model), the `reset` it issues is "not permitted when stream
    is capturing" and invalidates the capture in progress. Collect first, keep the collector off until the capture ends."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


class _Pool:
    """Static buffer assignment: exact-size free lists, so temporaries of equal shape share memory."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.total_bytes = 0

    def get(self, shape, dtype):
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self.total_bytes += t.numel() * t.element_size()
        return t

    def put(self, t):
        self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)


class _Act:
    """A residual-stream tensor (fp32 NHWC) and its GroupNorm statistics (parts [B, smax, C, 2] fp32, counts [B])."""
    __slots__ = ('t', 'stats')

    def __init__(self, t, stats):
        self.t, self.stats = t, stats


class UNetEngine:
    def __init__(self, model, batch, device):
        self.model = model
        self.B = batch
        self.device = device
        # fp32-class mode: kernels run with fp16 operands, every operand is a (hi, lo) pair and every GEMM has three K segments
        # per logical segment (hi hi + lo hi + hi lo); a "16-bit operand" below is then a tuple of two tensors
        self.split = model.operand_format == L.FMT_FP32X
        self.fmt = L.FMT_FP16 if self.split else model.operand_format
        self.wscale = L.SPLIT_WEIGHT_SCALE if self.split else 1.0
        self.dt16 = ops.dtype16(self.fmt)
        self.steps = []          # list of zero-arg callables, one kernel launch each
        self.n_embed_steps = 0   # leading steps (embedding front end) that run on a side stream ...
        self.join_idx = None     # ... until the first step that consumes the step embedding
        self.side = None
        self.early = []          # (launch, event): skip-half norms of the up path, run on a second side stream from ...
        self.early_fork_idx = None   # ... the first step of the middle block (the low-resolution levels leave HBM idle)
        self.side2 = None
        self.n_aux_steps = 0     # entries of self.steps that are stream waits, not kernels
        self.serial = False      # True: no side streams (per-kernel timing passes want kernels that do not overlap)
        self.packers = []        # callables that (re)fill the weight arena from the live parameters
        self.graph = None
        self.step_graphs = {}    # per-step CUDA graphs of the sampling loop (diffusion._StepRunner): they point into this
                                 # engine's buffers, so they live and die with it
        self.sig = None
        self.pool = _Pool(device)
        self._layouts = {}
        self._ones = torch.ones((batch,), dtype=torch.int32, device=device)
        with torch.no_grad():
            self._build()

    # ------------------------------------------------------------------------------------------
    # weight arena
    # ------------------------------------------------------------------------------------------
    # Packers look the parameter up BY NAME on its module every time they run, so a Parameter object that was replaced
    # (load_state_dict(assign=True), module surgery) is picked up by the next refresh.
    def _copy_param(self, mod, name, dtype=torch.float32):
        buf = torch.empty(getattr(mod, name).shape, dtype=dtype, device=self.device)
        self.packers.append(lambda: buf.copy_(getattr(mod, name).detach()))
        return buf

    def _packed(self, mod, name, cout_pad=None, cin_off=0, cin_len=None):
        p = getattr(mod, name)
        cout = p.shape[0]
        k = p.shape[2] if p.dim() >= 3 else 1
        cp = ops.pad16(cout) if cout_pad is None else cout_pad
        cl = p.shape[1] - cin_off if cin_len is None else cin_len
        bufs = [torch.empty((k * k, cp, cl), dtype=self.dt16, device=self.device) for _ in range(2 if self.split else 1)]

        def pack():
            w = getattr(mod, name).detach().to(self.device)
            for part, buf in enumerate(bufs):
                ops.pack_conv_weight(w, cout_pad=cp, cin_off=cin_off, cin_len=cl, fmt=self.fmt, out=buf, scale=self.wscale,
                                     part=part)
        self.packers.append(pack)
        return tuple(bufs) if self.split else bufs[0]

    # ---- 16-bit operands: one tensor, or a (hi, lo) pair in the fp32-class mode --------------------------------------
    def _new16(self, shape, pooled=True):
        mk = (lambda: self.pool.get(shape, self.dt16)) if pooled else \
            (lambda: torch.empty(shape, dtype=self.dt16, device=self.device))
        return (mk(), mk()) if self.split else mk()

    def _free16(self, t):
        if t is None:
            return
        for x in (t if isinstance(t, tuple) else (t,)):
            self.pool.put(x)

    def _expand(self, segs):
        """Logical K segments -> kernel segments (+ the accumulator scale that undoes the weight scale)."""
        if not self.split:
            return list(segs), 1.0
        out = []
        for a, w in segs:
            out += [(a[0], w[0]), (a[1], w[0]), (a[0], w[1])]
        if len(out) > L.MAX_SEGMENTS:
            raise L.NdbError('fp32-class mode: this layer needs {} K segments (max {})'.format(len(out), L.MAX_SEGMENTS))
        return out, 1.0 / self.wscale

    def _gn(self, x0, stats0, gamma, beta, out, raw16=None, **kw):
        """ops.gn_apply with the output operand(s) given as tensors or (hi, lo) pairs."""
        if self.split:
            return ops.gn_apply(x0, stats0, gamma, beta, out=out[0], out_lo=out[1],
                                raw16=raw16[0] if raw16 is not None else None,
                                raw_lo=raw16[1] if raw16 is not None else None, fmt=self.fmt, **kw)
        return ops.gn_apply(x0, stats0, gamma, beta, out=out, raw16=raw16, fmt=self.fmt, **kw)

    def _bias_sum(self, m1, m2):
        buf = torch.empty(m1.bias.shape, dtype=torch.float32, device=self.device)
        self.packers.append(lambda: torch.add(m1.bias.detach(), m2.bias.detach(), out=buf))
        return buf

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def refresh_weights(self, force=False):
        """Re-pack the weight arena when any parameter's storage or version changed. In-place writes through
        `param.data` (EMA copies, `p.data.mul_()`) bump no version counter: callers that may have done that pass
        force=True -- `Diffusion.denoise` does so at the start of every chain (a few ms, once per chain)."""
        sig = self._signature()
        if force or sig != self.sig:
            with torch.no_grad():
                for f in self.packers:
                    f()
            self.sig = sig

    # ------------------------------------------------------------------------------------------
    # emit helpers
    # ------------------------------------------------------------------------------------------
    def _stats_buf(self, C, smax=1, counts=None):
        parts = self.pool.get((self.B, smax, C, 2), torch.float32)
        if counts is None:
            counts = self._ones
        return (parts, counts)

    def _release(self, act):
        self.pool.put(act.t)
        if act.stats is not None:
            self.pool.put(act.stats[0])

    def _emit_stats_kernel(self, t, stats):
        self.steps.append(lambda: ops.gn_stats(t, stats0=stats, workspace=self.gn_ws))

    def _conv_stats_layout(self, H, W, cout, k_total):
        key = (H, W, cout, k_total)
        if key not in self._layouts:
            import ctypes
            counts = (ctypes.c_int32 * self.B)()
            smax = int(L.load().ndb_conv_stats_layout(self.B, H, W, ops.pad16(cout), 0, k_total,
                                                      L.SPLIT_ACC_CHUNK if self.split else 0, counts))
            if smax < 1:
                raise L.NdbError('ndb_conv_stats_layout failed')
            self._layouts[key] = (smax, torch.tensor(list(counts), dtype=torch.int32, device=self.device))
        return self._layouts[key]

    def _conv(self, segs, H, W, cout, want_stats, out=None, **kw):
        """Emit one implicit-GEMM launch. With want_stats the epilogue also produces the output's per-channel
        statistics (or, for geometries it does not cover, a standalone statistics launch follows)."""
        B, fmt = self.B, self.fmt
        segs, osc = self._expand(segs)
        if self.split:
            kw['out_scale'] = osc
            kw['acc_chunk'] = L.SPLIT_ACC_CHUNK
        if out is None:
            out = self.pool.get((B, H, W, cout), torch.float32)
        stats = None
        fused = want_stats and ops.conv_stats_supported(B, H, W)
        if fused:
            k_total = sum(a.shape[-1] * w.shape[0] for a, w in segs)
            smax, counts = self._conv_stats_layout(H, W, cout, k_total)
            stats = self._stats_buf(cout, smax, counts)
            self.steps.append(lambda: ops.conv_gemm(segs, B, H, W, cout, out=out, fmt=fmt, stats=stats, **kw))
        else:
            self.steps.append(lambda: ops.conv_gemm(segs, B, H, W, cout, out=out, fmt=fmt, **kw))
            if want_stats:
                stats = self._stats_buf(cout)
                self._emit_stats_kernel(out, stats)
        return _Act(out, stats)

    # ------------------------------------------------------------------------------------------
    # plan construction
    # ------------------------------------------------------------------------------------------
    def _build(self):
        m, B, dev, fmt = self.model, self.B, self.device, self.fmt
        R = m.resolution
        E = 4 * m.model_channels
        for c in self._all_channels():
            if c % 32 != 0:        # GroupNorm(32); a last K block of 32 channels is zero-filled by TMA
                raise L.NdbError('ndb200 needs every feature-map width to be a multiple of 32 (got {})'.format(c))

        # static inputs / output
        self.x_in = torch.zeros((B, m.in_channels, R, R), dtype=torch.float32, device=dev)
        self.t_in = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.y_in = torch.zeros((B,), dtype=torch.int64, device=dev) if m.conditional else None
        self.out = torch.empty((B, m.out_channels, R, R), dtype=torch.float32, device=dev)

        # ---- embedding front end + one GEMM for every ResBlock's step_embedding Linear ----
        freqs = timestep_embedding_freqs(m.model_channels).to(dev)
        w0 = self._copy_param(m.step_embed[0], 'weight'); b0 = self._copy_param(m.step_embed[0], 'bias')
        w2 = self._copy_param(m.step_embed[2], 'weight'); b2 = self._copy_param(m.step_embed[2], 'bias')
        ce = self._copy_param(m.class_embedding, 'weight') if m.conditional else None
        act16 = torch.empty((B, E), dtype=self.dt16, device=dev)
        self.emb = torch.empty((B, E), dtype=torch.float32, device=dev)
        h1 = torch.empty((B, E), dtype=torch.float32, device=dev)
        self.steps.append(lambda: ops.embed(self.t_in, self.y_in, freqs, w0, b0, w2, b2, ce, emb_out=self.emb,
                                            act_out=act16, fmt=fmt, scratch=h1))
        resblocks = [mod for mod in m.modules() if isinstance(mod, ResidualBlock)]
        total = sum(rb.step_embedding.out_features for rb in resblocks)
        assert total % 16 == 0
        w_parts = [torch.empty((1, total, E), dtype=self.dt16, device=dev) for _ in range(2 if self.split else 1)]
        b_all = torch.empty((total,), dtype=torch.float32, device=dev)
        self.emb_all = torch.empty((B, total), dtype=torch.float32, device=dev)
        self.emb_off = {}
        off = 0
        for rb in resblocks:
            n = rb.step_embedding.out_features
            self.emb_off[id(rb)] = off
            bv, lin = b_all[off:off + n], rb.step_embedding

            def pack_lin(lin=lin, n=n, off=off):
                w = lin.weight.detach().float() * self.wscale
                hi = w.to(self.dt16)
                w_parts[0][:, off:off + n, :].copy_(hi.view(1, n, E))
                if self.split:
                    w_parts[1][:, off:off + n, :].copy_((w - hi.float()).to(self.dt16).view(1, n, E))
            self.packers.append(pack_lin)
            self.packers.append(lambda bv=bv, lin=lin: bv.copy_(lin.bias.detach()))
            off += n
        if self.split:
            # SiLU(emb) as an fp16 (hi, lo) pair, from the fp32 embedding
            a_hi = torch.empty((B, E), dtype=self.dt16, device=dev)
            a_lo = torch.empty((B, E), dtype=self.dt16, device=dev)
            self.steps.append(lambda: ops.split16(self.emb, a_hi, a_lo, silu=True))
            a_view, w_all = (a_hi.view(1, 1, B, E), a_lo.view(1, 1, B, E)), tuple(w_parts)
        else:
            a_view, w_all = act16.view(1, 1, B, E), w_parts[0]
        esegs, eosc = self._expand([(a_view, w_all)])
        self.steps.append(lambda: ops.conv_gemm(esegs, 1, 1, B, total, bias=b_all, out=self.emb_all, fmt=fmt, cout_pad=total,
                                                out_scale=eosc, acc_chunk=L.SPLIT_ACC_CHUNK if self.split else 0))
        self.n_embed_steps = len(self.steps)

        # ---- workspace of the standalone statistics kernel (stem output, uncovered geometries) ----
        ws_bytes = 0
        res = R
        while res >= 1:
            ws_bytes = max(ws_bytes, int(L.load().ndb_gn_stats_workspace_bytes(B, res * res, 4096)))
            res //= 2
        self.gn_ws = torch.zeros((ws_bytes + 7) // 8, dtype=torch.int64, device=dev)

        # ---- trunk ----
        stem = m.downsampling[0][0]
        C0 = stem.weight.shape[0]
        if 9 * m.in_channels <= 64 and ops.conv_stats_supported(B, R, R) and os.environ.get('NDB_STEM_GEMM') and not self.split:
            # stem conv on the tensor cores: the 3x3 neighbourhood of a pixel (9 * Cin <= 64 values) is ONE K block, so the
            # conv is a 1x1 GEMM that is purely output-bandwidth bound and gets the fused bias / TMA-store epilogue /
            # GroupNorm statistics of every other conv. OPT-IN (NDB_STEM_GEMM=1): measured 0.05 ms (0.5 %) per step faster
            # than the direct fp32 kernel + statistics pass, but rounding the model input to 16 bits costs 8-12 % of the
            # error budget (ImageNet-64 forward 8.2e-4 -> 8.9e-4), so the exact fp32 stem stays the default
            cols = torch.empty((B, R, R, 64), dtype=self.dt16, device=dev)
            stem_wp = torch.empty((1, ops.pad16(C0), 64), dtype=self.dt16, device=dev)
            stem_b = self._copy_param(stem, 'bias')
            self.packers.append(lambda: ops.pack_conv_weight(ops.stem_weight_as_1x1(stem.weight.detach().to(dev).float()),
                                                             cout_pad=ops.pad16(C0), fmt=fmt, out=stem_wp))
            self.steps.append(lambda: ops.stem_im2col(self.x_in, out=cols, fmt=fmt))
            x = self._conv([(cols, stem_wp)], R, R, C0, want_stats=True, bias=stem_b)
        else:
            stem_w = self._copy_param(stem, 'weight'); stem_b = self._copy_param(stem, 'bias')
            if (R * R) % 64 == 0 and R * R // 64 <= 64:
                # the stem kernel also emits the statistics of its output (one slot per 64-pixel block)
                x = _Act(self.pool.get((B, R, R, C0), torch.float32), ops.stem_stats_alloc(B, R, R, C0, dev))
                self.steps.append(lambda: ops.stem_conv(self.x_in, stem_w, stem_b, out=x.t, stats=x.stats))
            else:
                x = _Act(self.pool.get((B, R, R, C0), torch.float32), self._stats_buf(C0))
                self.steps.append(lambda: ops.stem_conv(self.x_in, stem_w, stem_b, out=x.t))
                self._emit_stats_kernel(x.t, x.stats)
        xs = [x]
        cur = x
        for seq in list(m.downsampling)[1:]:
            cur = self._run_sequential(seq, cur, None, input_saved=True)
            xs.append(cur)
        self.early_fork_idx = len(self.steps)
        cur = self._run_sequential(m.middle_block, cur, None, input_saved=True)
        for seq in m.upsampling:
            skip = xs.pop()
            new = self._run_sequential(seq, cur, skip, input_saved=False)
            self._release(skip)
            cur = new
        # head: GroupNorm -> SiLU -> conv3x3 to out_channels (model.py:446-449), NCHW fp32 out
        gn, conv = m.out[0], m.out[2]
        g = self._copy_param(gn, 'weight'); bt = self._copy_param(gn, 'bias')
        a = self._new16(tuple(cur.t.shape))
        wp = self._packed(conv, 'weight'); bp = self._copy_param(conv, 'bias')
        H = cur.t.shape[1]
        self.steps.append(lambda: self._gn(cur.t, cur.stats, g, bt, a, silu=True, eps=gn.eps))
        hsegs, hosc = self._expand([(a, wp)])
        self.steps.append(lambda: ops.conv_gemm(hsegs, B, H, H, m.out_channels, bias=bp, out_kind=L.OUT_F32_NCHW, out=self.out,
                                                fmt=fmt, out_scale=hosc, acc_chunk=L.SPLIT_ACC_CHUNK if self.split else 0))
        self.num_kernels = len(self.steps) - self.n_aux_steps + len(self.early) + 1   # the embedding front end is two launches

    def _all_channels(self):
        cs = set()
        for mod in self.model.modules():
            if isinstance(mod, ResidualBlock):
                cs.add(mod.in_channels); cs.add(mod.out_channels)
            elif isinstance(mod, AttentionBlock):
                cs.add(mod.channels)
        return cs

    def _run_sequential(self, seq, x, skip, input_saved):
        """Emit the launches of one UsesStepsSequential (model.py:40-48). `x` (+ optional concat partner `skip`
        for the first layer, model.py:474) is the input; returns the output. Buffers of consumed intermediates go
        back to the pool; `input_saved` says the input is still referenced by the skip stack."""
        cur, cur_saved, first = x, input_saved, True
        for layer in seq:
            if isinstance(layer, ResidualBlock):
                out = self._resblock(layer, cur, skip if first else None)
            elif isinstance(layer, AttentionBlock):
                out = self._attention(layer, cur)
            elif isinstance(layer, Upsample):
                out = self._upsample(layer, cur)
            elif isinstance(layer, Downsample):
                out = self._downsample(layer, cur)
            else:
                raise L.NdbError('unexpected layer {}'.format(type(layer)))
            if not cur_saved:
                self._release(cur)
            cur, cur_saved, first = out, False, False
        return cur

    # ---- blocks ---------------------------------------------------------------------------
    def _resblock(self, rb, x, skip):
        B, fmt, pool = self.B, self.fmt, self.pool
        _, H, W, C0 = x.t.shape
        C1 = skip.t.shape[-1] if skip is not None else 0
        Cin, Cout = rb.in_channels, rb.out_channels
        assert C0 + C1 == Cin, (C0, C1, Cin)
        mode = {'same': L.RESAMPLE_NONE, 'up': L.RESAMPLE_UP2, 'down': L.RESAMPLE_DOWN2}[rb.mode]
        Ho, Wo = (H, W) if rb.mode == 'same' else ((2 * H, 2 * W) if rb.mode == 'up' else (H // 2, W // 2))
        conv_skip = not isinstance(rb.skip, torch.nn.Identity)
        if not conv_skip and skip is not None:
            raise L.NdbError('identity skip over a concatenated input is not supported')

        g1 = self._copy_param(rb.in_norm, 'weight'); b1 = self._copy_param(rb.in_norm, 'bias')
        g2 = self._copy_param(rb.out_norm, 'weight'); b2 = self._copy_param(rb.out_norm, 'bias')
        bias_in = self._copy_param(rb.in_conv, 'bias')
        w_out = self._packed(rb.out_conv, 'weight')
        # Concatenated input [x ; skip] whose GroupNorm groups do not straddle the seam: the skip half depends only on the
        # down path, so its norm + SiLU + 16-bit cast run EARLY on a side stream (forked at the middle block, when the small
        # feature maps leave HBM bandwidth idle); the convs read the two halves as separate K segments.
        cpg = Cin // 32
        split = (skip is not None and rb.mode == 'same' and conv_skip and Cin % 32 == 0 and C0 % cpg == 0
                 and self.early_fork_idx is not None and not os.environ.get('NDB_DISABLE_SPLIT_NORM'))
        if split:
            w_in_a = self._packed(rb.in_conv, 'weight', cin_off=0, cin_len=C0)
            w_in_b = self._packed(rb.in_conv, 'weight', cin_off=C0, cin_len=C1)
            w_skip_a = self._packed(rb.skip, 'weight', cin_off=0, cin_len=C0)
            w_skip_b = self._packed(rb.skip, 'weight', cin_off=C0, cin_len=C1)
        else:
            w_in = self._packed(rb.in_conv, 'weight')
        if conv_skip:
            if not split:
                w_skip = self._packed(rb.skip, 'weight')
            bias_out = self._bias_sum(rb.out_conv, rb.skip)
        else:
            bias_out = self._copy_param(rb.out_conv, 'bias')
        off = self.emb_off[id(rb)]
        if rb.use_adaptive_gn:
            scale = self.emb_all[:, off:off + Cout]
            shift = self.emb_all[:, off + Cout:off + 2 * Cout]
            pre = None
        else:
            scale = shift = None
            pre = self.emb_all[:, off:off + Cout]

        eps1, eps2 = rb.in_norm.eps, rb.out_norm.eps
        a2 = self._new16((B, Ho, Wo, Cout))
        pooled = None
        if split:
            a1 = self._new16((B, H, W, C0))
            raw = self._new16((B, H, W, C0))
            a1b = self._new16((B, H, W, C1), pooled=False)     # live from the fork to this block
            rawb = self._new16((B, H, W, C1), pooled=False)
            g1a, b1a, g1b, b1b = g1[:C0], b1[:C0], g1[C0:], b1[C0:]
            sk = skip
            ev = torch.cuda.Event()
            self.early.append((lambda: self._gn(sk.t, sk.stats, g1b, b1b, a1b, raw16=rawb, silu=True, eps=eps1, cpg=cpg), ev))
            self.steps.append(lambda: self._gn(x.t, x.stats, g1a, b1a, a1, raw16=raw, silu=True, eps=eps1, cpg=cpg))
            self.steps.append(lambda: torch.cuda.current_stream(self.device).wait_event(ev))
            self.n_aux_steps += 1
            h = self._conv([(a1, w_in_a), (a1b, w_in_b)], Ho, Wo, Cout, want_stats=True, bias=bias_in)
        else:
            a1 = self._new16((B, Ho, Wo, Cin))
            raw = self._new16((B, Ho, Wo, Cin)) if conv_skip else None
            x1 = skip.t if skip is not None else None
            s1 = skip.stats if skip is not None else None
            # down blocks with an identity skip: the apply kernel (which reads x anyway) also leaves avg_pool2d(x) in fp32,
            # so the out_conv epilogue adds it as a plain residual instead of gathering four taps per element
            pooled = pool.get((B, Ho, Wo, Cin), torch.float32) if (rb.mode == 'down' and not conv_skip) else None
            self.steps.append(lambda: self._gn(x.t, x.stats, g1, b1, a1, raw16=raw, x1=x1, stats1=s1, silu=True, resample=mode,
                                               eps=eps1, pool32=pooled))
            if self.model.h_fp16 and not self.split and ops.conv_stats_supported(B, Ho, Wo):
                # `h` only feeds the out_norm: 16-bit storage halves its write + read (statistics still come from the fp32
                # accumulators in the epilogue)
                h = self._conv([(a1, w_in)], Ho, Wo, Cout, want_stats=True, bias=bias_in, out_kind=L.OUT_F16_NHWC,
                               out=pool.get((B, Ho, Wo, Cout), self.dt16))
            else:
                h = self._conv([(a1, w_in)], Ho, Wo, Cout, want_stats=True, bias=bias_in)
        if self.join_idx is None:
            self.join_idx = len(self.steps)     # first consumer of the step embedding
        self.steps.append(lambda: self._gn(h.t, h.stats, g2, b2, a2, scale=scale, shift=shift, pre_add=pre, silu=True,
                                           eps=eps2))
        if split:
            out = self._conv([(a2, w_out), (raw, w_skip_a), (rawb, w_skip_b)], Ho, Wo, Cout, want_stats=True, bias=bias_out)
        elif conv_skip:
            # 1x1 (or 3x3 when use_conv) skip conv as an extra K range of the same GEMM
            out = self._conv([(a2, w_out), (raw, w_skip)], Ho, Wo, Cout, want_stats=True, bias=bias_out)
        else:
            rmode = {'same': L.RES_SAME, 'up': L.RES_UP2, 'down': L.RES_SAME}[rb.mode]
            out = self._conv([(a2, w_out)], Ho, Wo, Cout, want_stats=True, bias=bias_out,
                             residual=pooled if pooled is not None else x.t, residual_mode=rmode)
        for t in (a1, raw, a2):
            self._free16(t)
        if pooled is not None:
            pool.put(pooled)
        self._release(h)
        return out

    def _attention(self, ab, x):
        B, fmt, pool = self.B, self.fmt, self.pool
        _, H, W, C = x.t.shape
        T = H * W
        heads = ab.num_heads
        d = C // heads
        g = self._copy_param(ab.norm, 'weight'); bt = self._copy_param(ab.norm, 'bias')
        w_qkv = self._packed(ab.qkv_nin, 'weight'); b_qkv = self._copy_param(ab.qkv_nin, 'bias')
        w_proj = self._packed(ab.proj_out, 'weight'); b_proj = self._copy_param(ab.proj_out, 'bias')
        n16 = self._new16((B, H, W, C))
        o16 = self._new16((B, H, W, C))
        order = 0 if ab.split_qkv_first else 1
        eps = ab.norm.eps
        self.steps.append(lambda: self._gn(x.t, x.stats, g, bt, n16, silu=False, eps=eps))
        if self.split:
            # fp32-class: qkv in fp32 -> (hi, lo) pairs -> three-product attention -> fp32 -> (hi, lo) pairs for proj
            qkv32 = pool.get((B, H, W, 3 * C), torch.float32)
            qkv = self._new16((B, H, W, 3 * C))
            o32 = pool.get((B, H, W, C), torch.float32)
            self._conv([(n16, w_qkv)], H, W, 3 * C, want_stats=False, out=qkv32, bias=b_qkv)
            self.steps.append(lambda: ops.split16(qkv32, qkv[0], qkv[1]))
            self.steps.append(lambda: ops.attention_split(qkv[0].view(B, T, 3 * C), qkv[1].view(B, T, 3 * C), B, T, heads, d,
                                                          order=order, out=o32.view(B, T, C)))
            self.steps.append(lambda: ops.split16(o32, o16[0], o16[1]))
            pool.put(qkv32); pool.put(o32)
        else:
            qkv = pool.get((B, H, W, 3 * C), self.dt16)
            self._conv([(n16, w_qkv)], H, W, 3 * C, want_stats=False, out=qkv, bias=b_qkv, out_kind=L.OUT_F16_NHWC)
            self.steps.append(lambda: ops.attention(qkv.view(B, T, 3 * C), B, T, heads, d, order=order,
                                                    out=o16.view(B, T, C), fmt=fmt))
        out = self._conv([(o16, w_proj)], H, W, C, want_stats=True, bias=b_proj, residual=x.t,
                         residual_mode=L.RES_SAME)
        for t in (n16, qkv, o16):
            self._free16(t)
        return out

    def _upsample(self, layer, x):
        """Standalone Upsample (resblock_updown=False): nearest x2 [+ conv3x3] (model.py:76-80)."""
        B, fmt, pool = self.B, self.fmt, self.pool
        _, H, W, C = x.t.shape
        up = pool.get((B, 2 * H, 2 * W, C), torch.float32)
        self.steps.append(lambda: ops.resample_f32(x.t, L.RESAMPLE_UP2, out=up))
        if not layer.with_conv:
            act = _Act(up, self._stats_buf(C))
            self._emit_stats_kernel(up, act.stats)
            return act
        Cout = layer.conv.weight.shape[0]
        a = self._new16((B, 2 * H, 2 * W, C))
        wp = self._packed(layer.conv, 'weight'); bp = self._copy_param(layer.conv, 'bias')
        if self.split:
            self.steps.append(lambda: ops.split16(up, a[0], a[1]))
        else:
            self.steps.append(lambda: ops.cast16(up, out=a, fmt=fmt))
        out = self._conv([(a, wp)], 2 * H, 2 * W, Cout, want_stats=True, bias=bp)
        pool.put(up); self._free16(a)
        return out

    def _downsample(self, layer, x):
        """Standalone Downsample: conv3x3 stride 2 (as nine 1x1 K-segments over a gathered operand) or avg-pool."""
        B, fmt, pool = self.B, self.fmt, self.pool
        _, H, W, C = x.t.shape
        if not layer.with_conv:
            out = pool.get((B, H // 2, W // 2, C), torch.float32)
            self.steps.append(lambda: ops.resample_f32(x.t, L.RESAMPLE_DOWN2, out=out))
            act = _Act(out, self._stats_buf(C))
            self._emit_stats_kernel(out, act.stats)
            return act
        if self.split:
            raise L.NdbError('fp32-class mode: the strided Downsample conv (nine K segments, 27 with split operands) is not '
                             'supported; use resblock_updown=True or conv_resample=False')
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        Cout = layer.conv.weight.shape[0]
        cols = pool.get((9, B, Ho, Wo, C), self.dt16)
        wp = self._packed(layer.conv, 'weight'); bp = self._copy_param(layer.conv, 'bias')
        segs = [(cols[tap], wp[tap:tap + 1]) for tap in range(9)]
        self.steps.append(lambda: ops.im2col_s2(x.t, out=cols, fmt=fmt))
        out = self._conv(segs, Ho, Wo, Cout, want_stats=True, bias=bp)
        pool.put(cols)
        return out

    # ------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------
    def _launch_all(self):
        """The embedding front end (timestep / class embedding MLP + the GEMM of all step-embedding Linears: small grids,
        independent of x) runs on a side stream next to the stem conv, its statistics and the first norm + conv; the
        streams join at the first kernel that reads the embedding. Under graph capture this becomes a fork / join."""
        ne = self.n_embed_steps
        cur = torch.cuda.current_stream(self.device)
        if self.serial:
            for i, f in enumerate(self.steps):
                if self.early and i == self.early_fork_idx:
                    for g, ev in self.early:
                        g()
                        ev.record(cur)
                f()
            return
        fork_embed = ne > 0 and self.join_idx is not None
        if fork_embed:
            if self.side is None:
                self.side = torch.cuda.Stream(device=self.device)
            self.side.wait_stream(cur)
            with torch.cuda.stream(self.side):
                for f in self.steps[:ne]:
                    f()
        for i in range(ne if fork_embed else 0, len(self.steps)):
            if fork_embed and i == self.join_idx:
                cur.wait_stream(self.side)
            if self.early and i == self.early_fork_idx:
                # skip-half norms of the up path: everything they read exists by now; each records an event its consumer
                # waits on
                if self.side2 is None:
                    self.side2 = torch.cuda.Stream(device=self.device)
                self.side2.wait_stream(cur)
                with torch.cuda.stream(self.side2):
                    for f, ev in self.early:
                        f()
                        ev.record(self.side2)
            self.steps[i]()

    def run_static(self):
        """Run the plan on the static input buffers (x_in, t_in, y_in) -> self.out."""
        if self.model.use_cuda_graph:
            if self.graph is None:
                # warm up eagerly on a side stream (also sets kernel attributes), then capture
                s = torch.cuda.Stream(device=self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._launch_all()
                torch.cuda.current_stream(self.device).wait_stream(s)
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with capture_guard(), torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
            self.graph.replay()
            ops.LAUNCHES += self.num_kernels
        else:
            self._launch_all()
        return self.out

    @torch.no_grad()
    def timeline(self, replays=5):
        """In-graph, warm-cache duration of every launch of the plan: the launches are captured on ONE stream with a
        %globaltimer stamp (ndb_timestamp) between consecutive kernels, the graph is replayed and the stamp differences
        averaged. Returns ([(kernel name, description, microseconds, algorithmic flops, algorithmic bytes)], stamp cost),
        with the stamp kernel's own cost (two stamps back to back) subtracted. Unlike CUDA events around eager launches this does not include host launch latency, and
        unlike an ncu pass the caches are as warm as in production."""
        dev = self.device
        launches = []                 # (callable, is_kernel)
        for i, f in enumerate(self.steps):
            if self.early and i == self.early_fork_idx:
                launches += [(g, ev) for g, ev in self.early]    # their consumers wait on these events
            launches.append((f, None))
        n = len(launches)
        buf = torch.zeros(n + 3, dtype=torch.int64, device=dev)
        was_serial, self.serial = self.serial, True
        names = []

        def run(record):
            ops.timestamp(buf, 0)
            ops.timestamp(buf, 1)                      # calibration: stamp[1] - stamp[0] = cost of a stamp
            for k, (f, ev) in enumerate(launches):
                n0 = len(ops.PROFILE) if record else 0
                f()
                if ev is not None:
                    ev.record(torch.cuda.current_stream(dev))
                if record:
                    rec = ops.PROFILE[n0:]
                    names.append((rec[0][0], rec[0][4], rec[0][3], rec[0][5]) if rec else ('wait', '', 0, 0))
                ops.timestamp(buf, 2 + k)
        try:
            ops.PROFILE = []
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                run(True)
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            ops.PROFILE = None
            g = torch.cuda.CUDAGraph()
            with capture_guard(), torch.cuda.graph(g):
                run(False)
            acc = torch.zeros(n + 3, dtype=torch.float64)
            for _ in range(replays):
                g.replay()
                torch.cuda.synchronize(dev)
                acc += buf.cpu().double()
            acc /= replays
        finally:
            ops.PROFILE = None
            self.serial = was_serial
        stamp_cost = float(acc[1] - acc[0])
        out = []
        for k in range(n):
            out.append((names[k][0], names[k][1], max(float(acc[2 + k] - acc[1 + k]) - stamp_cost, 0.0) / 1e3,
                        names[k][2], names[k][3]))
        return out, stamp_cost / 1e3

    @torch.no_grad()
    def forward(self, x, timestep, y):
        self.refresh_weights()
        self.x_in.copy_(x)
        self.t_in.copy_(timestep)
        if self.y_in is not None:
            self.y_in.copy_(y)
        return self.run_static().clone()
