"""Tensor-level wrappers of the ndb200 kernels: allocate outputs with torch (device memory + stream plumbing
only), pass raw pointers through ctypes. Every function launches exactly one CUDA kernel of this library
on the current stream; `LAUNCHES` counts them (bench.py reports it as gpu_launches).

Layouts: activations NHWC. `*16` tensors are 16-bit tensor-core operands (torch.float16 by default,
torch.bfloat16 when fmt == FMT_BF16).
"""
import ctypes as C

import torch

from . import _lib as L

LAUNCHES = 0
PROFILE = None   # set to a list to record (kernel name, start event, end event, algorithmic flops) per launch
_pending = None


def _begin():
    global _pending
    if PROFILE is not None:
        _pending = torch.cuda.Event(enable_timing=True)
        _pending.record()


def _count(name='kernel', flops=0, desc='', nbytes=0):
    global LAUNCHES
    LAUNCHES += 1
    if PROFILE is not None:
        end = torch.cuda.Event(enable_timing=True)
        end.record()
        PROFILE.append((name, _pending, end, flops, desc, nbytes))


def profile_summary(records):
    """Aggregate PROFILE records (after a device synchronize) -> {name: {ms, launches, flops}, total_ms}."""
    out, total = {}, 0.0
    for name, e0, e1, flops, _desc, _nbytes in records:
        ms = e0.elapsed_time(e1)
        d = out.setdefault(name, {'ms': 0.0, 'launches': 0, 'flops': 0})
        d['ms'] += ms
        d['launches'] += 1
        d['flops'] += flops
        total += ms
    out['total_ms'] = total
    return out


def dtype16(fmt):
    return torch.float16 if fmt == L.FMT_FP16 else torch.bfloat16


def pad16(n):
    return (n + 15) // 16 * 16


def pack_conv_weight(w, cout_pad=None, cin_off=0, cin_len=None, fmt=L.FMT_FP16, out=None, scale=1.0, part=0):
    """nn.Conv2d / Conv1d / Linear weight (fp32, [cout, cin, k, k] | [cout, cin, 1] | [cout, cin]) ->
    [taps][cout_pad][cin_len] 16-bit rounding of w * scale (part 0) or the fp16 remainder of that rounding (part 1)."""
    L.require_device(w, 'pack_conv_weight')
    w = w.detach().contiguous().float()
    cout, cin = w.shape[0], w.shape[1]
    k = w.shape[2] if w.dim() >= 3 else 1
    if w.dim() == 4:
        assert w.shape[2] == w.shape[3]
    cout_pad = pad16(cout) if cout_pad is None else cout_pad
    cin_len = cin - cin_off if cin_len is None else cin_len
    if out is None:
        out = torch.empty((k * k, cout_pad, cin_len), dtype=dtype16(fmt), device=w.device)
    _begin()
    L.check(L.load().ndb_pack_conv_weight(L.ptr(w), L.ptr(out), cout, cin, k, cout_pad, cin_off, cin_len, fmt,
                                          float(scale), int(part), L.stream_ptr()), 'ndb_pack_conv_weight')
    _count('pack_conv_weight')
    return out


def conv_gemm(segments, N, H, W, cout, bias=None, residual=None, residual_mode=L.RES_NONE,
              out_kind=L.OUT_F32_NHWC, out=None, fmt=L.FMT_FP16, block_n=0, cout_pad=None,
              stats=None, out_scale=1.0, acc_chunk=0):
    """segments: list of (act16 [N,H,W,C], packed_weight [taps,cout_pad,C]).  Returns `out`.
    out_scale: power of two applied to the accumulators before bias / residual (fp32-class mode: 1 / SPLIT_WEIGHT_SCALE)."""
    p = L.ConvParams()
    p.N, p.H, p.W = N, H, W
    p.cout = cout
    p.cout_pad = segments[0][1].shape[1] if cout_pad is None else cout_pad
    p.num_segments = len(segments)
    assert 1 <= len(segments) <= L.MAX_SEGMENTS
    dev = segments[0][0].device
    for i, (a, w) in enumerate(segments):
        L.require_device(a, 'conv_gemm')
        assert a.dtype == dtype16(fmt) and w.dtype == dtype16(fmt)
        assert a.is_contiguous() and w.is_contiguous()
        assert w.shape[1] == p.cout_pad and w.shape[2] == a.shape[-1], (w.shape, a.shape)
        p.seg[i].act = a.data_ptr()
        p.seg[i].weight = w.data_ptr()
        p.seg[i].C = a.shape[-1]
        p.seg[i].taps = w.shape[0]
    p.bias = bias.data_ptr() if bias is not None else None
    p.residual = residual.data_ptr() if residual is not None else None
    p.residual_mode = residual_mode if residual is not None else L.RES_NONE
    p.out_kind = out_kind
    if out is None:
        if out_kind == L.OUT_F32_NHWC:
            out = torch.empty((N, H, W, cout), dtype=torch.float32, device=dev)
        elif out_kind == L.OUT_F16_NHWC:
            out = torch.empty((N, H, W, cout), dtype=dtype16(fmt), device=dev)
        else:
            out = torch.empty((N, cout, H, W), dtype=torch.float32, device=dev)
    p.out = out.data_ptr()
    p.operand_format = fmt
    p.block_n = block_n
    p.out_scale = float(out_scale)
    p.acc_chunk_kblocks = int(acc_chunk)
    if stats is not None:
        parts = stats[0]   # float32 [N, smax, cout, 2]
        assert parts.dtype == torch.float32 and parts.shape[0] == N and parts.shape[2] == cout and parts.shape[3] == 2
        p.stat_parts = parts.data_ptr()
        p.stat_smax = parts.shape[1]
    _begin()
    L.check(L.load().ndb_conv_gemm(C.byref(p), L.stream_ptr()), 'ndb_conv_gemm')
    _count('conv_gemm', 2 * N * H * W * cout * sum(a.shape[-1] * w.shape[0] for a, w in segments),
           'N{} {}x{} K{} cout{} out{} res{}'.format(N, H, W, sum(a.shape[-1] * w.shape[0] for a, w in segments), cout,
                                                    out_kind, p.residual_mode))
    return out


def conv_stats_supported(N, H, W):
    return bool(L.load().ndb_conv_stats_supported(N, H, W))


def conv_stats_alloc(N, H, W, cout, device, k_total, cout_pad=None, block_n=0, acc_chunk=0):
    """Buffers for the statistics a conv epilogue writes for its output: (parts float32 [N, smax, cout, 2],
    counts int32 [N]) -- see ndb_conv_stats_layout. k_total = reduction length of the GEMM that will write them
    (sum over its segments of taps * C): it selects the kernel variant and with it the slot layout."""
    import ctypes
    cp = pad16(cout) if cout_pad is None else cout_pad
    counts = (ctypes.c_int32 * N)()
    smax = int(L.load().ndb_conv_stats_layout(N, H, W, cp, block_n, int(k_total), int(acc_chunk), counts))
    if smax < 1:
        raise L.NdbError('ndb_conv_stats_layout failed ({})'.format(smax))
    parts = torch.zeros((N, smax, cout, 2), dtype=torch.float32, device=device)
    return parts, torch.tensor(list(counts), dtype=torch.int32, device=device)


def gn_workspace(N, HW, C, device):
    """Zero-filled scratch for gn_stats (see ndb_gn_stats_workspace_bytes)."""
    nbytes = int(L.load().ndb_gn_stats_workspace_bytes(N, HW, C))
    return torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)


def gn_stats(x0, x1=None, stats0=None, stats1=None, workspace=None):
    """Standalone statistics producer: per-channel (sum, sumsq) over the pixels of each image, one slot per image:
    x0 [N,H,W,C0] fp32 -> (parts float32 [N,1,C0,2], counts int32 [N] of ones); same for an optional second tensor
    in the same launch. Deterministic (no fp atomics)."""
    L.require_device(x0, 'gn_stats')
    N = x0.shape[0]
    HW = x0.shape[1] * x0.shape[2]
    C0 = x0.shape[-1]
    C1 = x1.shape[-1] if x1 is not None else 0

    def fresh(C):
        return (torch.empty((N, 1, C, 2), dtype=torch.float32, device=x0.device),
                torch.ones((N,), dtype=torch.int32, device=x0.device))
    if stats0 is None:
        stats0 = fresh(C0)
    if x1 is not None and stats1 is None:
        stats1 = fresh(C1)
    if workspace is None:
        workspace = gn_workspace(N, HW, C0 + C1, x0.device)
    assert stats0[0].shape[1] == 1
    _begin()
    L.check(L.load().ndb_gn_stats(L.ptr(x0), C0, L.ptr(x1), C1, N, HW, L.ptr(stats0[0]),
                                  L.ptr(stats1[0]) if x1 is not None else None,
                                  L.ptr(workspace), L.stream_ptr()), 'ndb_gn_stats')
    _count('gn_stats', 0, 'N{} HW{} C{}'.format(N, HW, C0 + C1))
    return (stats0, stats1) if x1 is not None else stats0


def stats_totals(stats):
    """(parts, counts) -> float64 [N, C, 2] totals (test helper; the kernels never materialise this)."""
    parts, counts = stats
    N, smax = parts.shape[0], parts.shape[1]
    mask = torch.arange(smax, device=parts.device)[None, :] < counts[:, None]
    return torch.where(mask[:, :, None, None], parts.double(), torch.zeros((), dtype=torch.float64, device=parts.device)).sum(1)


def gn_apply(x0, stats0, gamma, beta, x1=None, stats1=None, scale=None, shift=None, pre_add=None, silu=True,
             resample=L.RESAMPLE_NONE, out=None, raw16=None, want_raw16=False, fmt=L.FMT_FP16, eps=1e-5, pool32=None, cpg=0,
             out_lo=None, raw_lo=None):
    """out_lo / raw_lo (fp32-class mode): tensors like out / raw16 that receive the fp16 remainders."""
    L.require_device(x0, 'gn_apply')
    N, H, W, C0 = x0.shape
    C1 = x1.shape[-1] if x1 is not None else 0
    Ho, Wo = H, W
    if resample == L.RESAMPLE_UP2:
        Ho, Wo = 2 * H, 2 * W
    elif resample == L.RESAMPLE_DOWN2:
        Ho, Wo = H // 2, W // 2
    if out is None:
        out = torch.empty((N, Ho, Wo, C0 + C1), dtype=dtype16(fmt), device=x0.device)
    if want_raw16 and raw16 is None:
        raw16 = torch.empty_like(out)
    p = L.GnApplyParams()
    p.x0, p.C0 = x0.data_ptr(), C0
    p.x0_is16 = 1 if x0.dtype != torch.float32 else 0
    assert x0.dtype in (torch.float32, dtype16(fmt))
    p.x1, p.C1 = (x1.data_ptr() if x1 is not None else None), C1
    p.N, p.H, p.W = N, H, W
    p.parts0, p.counts0, p.smax0 = stats0[0].data_ptr(), stats0[1].data_ptr(), stats0[0].shape[1]
    assert stats0[0].shape[2] == C0
    if x1 is not None:
        assert stats1 is not None and stats1[0].shape[2] == C1
        p.parts1, p.counts1, p.smax1 = stats1[0].data_ptr(), stats1[1].data_ptr(), stats1[0].shape[1]
    p.gamma, p.beta = gamma.data_ptr(), beta.data_ptr()
    if scale is not None:
        assert scale.stride(0) == shift.stride(0) and scale.stride(1) == 1
        p.scale, p.shift, p.ld_ss = scale.data_ptr(), shift.data_ptr(), scale.stride(0)
    if pre_add is not None:
        p.pre_add, p.ld_pre = pre_add.data_ptr(), pre_add.stride(0)
    p.silu = 1 if silu else 0
    p.resample = resample
    p.out = out.data_ptr()
    p.raw16 = raw16.data_ptr() if raw16 is not None else None
    p.out_lo = out_lo.data_ptr() if out_lo is not None else None
    p.raw_lo = raw_lo.data_ptr() if raw_lo is not None else None
    p.pool32 = pool32.data_ptr() if pool32 is not None else None
    p.operand_format = fmt
    p.eps = eps
    p.cpg = cpg
    _begin()
    L.check(L.load().ndb_gn_apply(C.byref(p), L.stream_ptr()), 'ndb_gn_apply')
    # algorithmic bytes: the input read once + every output written once
    nbytes = N * H * W * (C0 * x0.element_size() + C1 * 4) + N * Ho * Wo * (C0 + C1) * (
        2 + (2 if raw16 is not None else 0) + (4 if pool32 is not None else 0))
    _count('gn_apply', 0, 'N{} {}x{} C{} rs{} raw{}'.format(N, H, W, C0 + C1, resample, int(raw16 is not None)), nbytes)
    return (out, raw16) if (want_raw16 or raw16 is not None) else out


def attention(qkv16, B, T, heads, d, order=0, out=None, fmt=L.FMT_FP16):
    """qkv16 [B,T,3C] -> [B,T,C]."""
    L.require_device(qkv16, 'attention')
    assert qkv16.is_contiguous() and qkv16.dtype == dtype16(fmt)
    if out is None:
        out = torch.empty((B, T, heads * d), dtype=qkv16.dtype, device=qkv16.device)
    _begin()
    L.check(L.load().ndb_attention(L.ptr(qkv16), L.ptr(out), B, T, heads, d, order, fmt, L.stream_ptr()),
            'ndb_attention')
    _count('attention', 4 * B * heads * T * T * d, 'B{} T{} h{} d{}'.format(B, T, heads, d))
    return out


def attention_split(qkv_hi, qkv_lo, B, T, heads, d, order=0, out=None):
    """fp32-class attention: qkv as fp16 (hi, lo) pairs [B,T,3C] -> fp32 [B,T,C]."""
    L.require_device(qkv_hi, 'attention_split')
    assert qkv_hi.is_contiguous() and qkv_lo.is_contiguous() and qkv_hi.dtype == torch.float16 and qkv_lo.dtype == torch.float16
    if out is None:
        out = torch.empty((B, T, heads * d), dtype=torch.float32, device=qkv_hi.device)
    _begin()
    L.check(L.load().ndb_attention_split(L.ptr(qkv_hi), L.ptr(qkv_lo), L.ptr(out), B, T, heads, d, order, L.stream_ptr()),
            'ndb_attention_split')
    _count('attention', 3 * 4 * B * heads * T * T * d, 'B{} T{} h{} d{} split'.format(B, T, heads, d))
    return out


def embed(t, y, freqs, w0, b0, w2, b2, class_emb, emb_out=None, act_out=None, fmt=L.FMT_FP16, scratch=None):
    L.require_device(t, 'embed')
    B = t.shape[0]
    mc = w0.shape[1]
    E = 4 * mc
    assert t.dtype == torch.float32 and (y is None or y.dtype == torch.int64)
    if emb_out is None:
        emb_out = torch.empty((B, E), dtype=torch.float32, device=t.device)
    if act_out is None:
        act_out = torch.empty((B, E), dtype=dtype16(fmt), device=t.device)
    if scratch is None:
        scratch = torch.empty((B, E), dtype=torch.float32, device=t.device)
    _begin()
    L.check(L.load().ndb_embed(L.ptr(t), L.ptr(y), L.ptr(freqs), B, mc, L.ptr(w0), L.ptr(b0), L.ptr(w2), L.ptr(b2),
                               L.ptr(class_emb) if y is not None else None, L.ptr(scratch), L.ptr(emb_out), L.ptr(act_out), fmt,
                               L.stream_ptr()), 'ndb_embed')
    _count('embed')
    global LAUNCHES
    LAUNCHES += 1   # two kernels
    return emb_out, act_out


def stem_conv(x_nchw, w, bias, out=None, stats=None):
    """Exact fp32 stem conv. stats = (parts float32 [N, H*W/64, Cout, 2], counts int32 [N]) from stem_stats_alloc: the
    kernel also writes the output's GroupNorm statistics (one slot per 64-pixel block)."""
    L.require_device(x_nchw, 'stem_conv')
    N, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    if out is None:
        out = torch.empty((N, H, W, Cout), dtype=torch.float32, device=x_nchw.device)
    sp = None
    if stats is not None:
        assert (H * W) % 64 == 0 and tuple(stats[0].shape) == (N, H * W // 64, Cout, 2)
        sp = stats[0].data_ptr()
    _begin()
    L.check(L.load().ndb_stem_conv(L.ptr(x_nchw), L.ptr(w), L.ptr(bias), L.ptr(out), sp, N, Cin, H, W, Cout,
                                   L.stream_ptr()), 'ndb_stem_conv')
    _count('stem_conv')
    return out


def stem_stats_alloc(N, H, W, cout, device):
    """Statistics buffers in the layout ndb_stem_conv writes: every one of the H*W/64 slots of an image is used."""
    assert (H * W) % 64 == 0
    smax = H * W // 64
    return (torch.zeros((N, smax, cout, 2), dtype=torch.float32, device=device),
            torch.full((N,), smax, dtype=torch.int32, device=device))


def stem_im2col(x_nchw, out=None, fmt=L.FMT_FP16):
    """[N,Cin,H,W] fp32 -> [N,H,W,64] 16-bit rows (tap*Cin + c, zero padded): the stem conv's GEMM operand."""
    L.require_device(x_nchw, 'stem_im2col')
    N, Cin, H, W = x_nchw.shape
    if out is None:
        out = torch.empty((N, H, W, 64), dtype=dtype16(fmt), device=x_nchw.device)
    _begin()
    L.check(L.load().ndb_stem_im2col(L.ptr(x_nchw), L.ptr(out), N, Cin, H, W, fmt, L.stream_ptr()), 'ndb_stem_im2col')
    _count('stem_im2col')
    return out


def stem_weight_as_1x1(w):
    """OIHW [Cout,Cin,3,3] -> [Cout,64,1,1] with column tap*Cin + c (the layout ndb_stem_im2col produces), zero padded."""
    Cout, Cin = w.shape[0], w.shape[1]
    assert w.shape[2] == 3 and w.shape[3] == 3 and 9 * Cin <= 64
    w2 = torch.zeros((Cout, 64, 1, 1), dtype=w.dtype, device=w.device)
    w2[:, :9 * Cin, 0, 0] = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)
    return w2


def resample_f32(x, mode, out=None):
    N, H, W, Cc = x.shape
    Ho, Wo = (2 * H, 2 * W) if mode == L.RESAMPLE_UP2 else (H // 2, W // 2)
    if out is None:
        out = torch.empty((N, Ho, Wo, Cc), dtype=torch.float32, device=x.device)
    _begin()
    L.check(L.load().ndb_resample_f32(L.ptr(x), L.ptr(out), N, H, W, Cc, mode, L.stream_ptr()), 'ndb_resample_f32')
    _count('resample_f32')
    return out


def cast16(x, out=None, fmt=L.FMT_FP16):
    if out is None:
        out = torch.empty(x.shape, dtype=dtype16(fmt), device=x.device)
    _begin()
    L.check(L.load().ndb_cast_nhwc16(L.ptr(x), L.ptr(out), x.numel(), fmt, L.stream_ptr()), 'ndb_cast_nhwc16')
    _count('cast_nhwc16')
    return out


def split16(x, hi=None, lo=None, silu=False):
    """fp32 -> fp16 pair (hi, lo) with hi + lo == x to ~2^-22 relative (fp32-class mode operands)."""
    L.require_device(x, 'split16')
    x = x.contiguous()
    if hi is None:
        hi = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    if lo is None:
        lo = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _begin()
    L.check(L.load().ndb_split16(L.ptr(x), L.ptr(hi), L.ptr(lo), x.numel(), 1 if silu else 0, L.stream_ptr()), 'ndb_split16')
    _count('split16')
    return hi, lo


def im2col_s2(x, out=None, fmt=L.FMT_FP16):
    N, H, W, Cc = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    if out is None:
        out = torch.empty((9, N, Ho, Wo, Cc), dtype=dtype16(fmt), device=x.device)
    _begin()
    L.check(L.load().ndb_im2col_s2(L.ptr(x), L.ptr(out), N, H, W, Cc, fmt, L.stream_ptr()), 'ndb_im2col_s2')
    _count('im2col_s2')
    return out


def sampler_step(x_t, model_out, t, tables, var_type, ddim=False, eta=0.0, null_out=None, cfg_strength=0.0,
                 grad=None, clf_strength=0.0, clip_x0=True, noise=None, seed=0, sample_ids=None,
                 x_prev=None, pred_x0=None, x_prev2=None):
    L.require_device(x_t, 'sampler_step')
    B, Cc, H, W = x_t.shape
    p = L.StepParams()
    assert x_t.is_contiguous() and model_out.is_contiguous() and t.dtype == torch.float32
    assert tables.dtype == torch.float32 and tables.shape[0] == L.TAB_COUNT and tables.is_contiguous()
    p.x_t, p.model_out = x_t.data_ptr(), model_out.data_ptr()
    p.null_out = null_out.data_ptr() if null_out is not None else None
    p.t, p.tables = t.data_ptr(), tables.data_ptr()
    p.R = tables.shape[1]
    p.B, p.C, p.HW, p.Cm = B, Cc, H * W, model_out.shape[1]
    p.var_type = var_type
    p.ddim = 1 if ddim else 0
    p.eta = float(eta)
    p.cfg_strength = float(cfg_strength)
    p.grad = grad.data_ptr() if grad is not None else None
    p.clf_strength = float(clf_strength)
    p.clip_x0 = 1 if clip_x0 else 0
    p.noise = noise.data_ptr() if noise is not None else None
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    p.sample_ids = sample_ids.data_ptr() if sample_ids is not None else None
    if x_prev is None:
        x_prev = torch.empty_like(x_t)
    p.x_prev = x_prev.data_ptr()
    p.pred_x0 = pred_x0.data_ptr() if pred_x0 is not None else None
    p.x_prev2 = x_prev2.data_ptr() if x_prev2 is not None else None
    _begin()
    L.check(L.load().ndb_sampler_step(C.byref(p), L.stream_ptr()), 'ndb_sampler_step')
    _count('sampler_step')
    return x_prev


def advance_step(t_step, t_model, timestep_map_f32):
    _begin()
    L.check(L.load().ndb_advance_step(L.ptr(t_step), t_step.shape[0], L.ptr(t_model), t_model.shape[0],
                                      L.ptr(timestep_map_f32), L.stream_ptr()), 'ndb_advance_step')
    _count('advance_step')


def philox_normal(shape, seed, sample_ids, step, device, out=None):
    B = shape[0]
    per = 1
    for s in shape[1:]:
        per *= s
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=device)
    _begin()
    L.check(L.load().ndb_philox_normal(L.ptr(out), B, per, int(seed) & 0xFFFFFFFFFFFFFFFF, L.ptr(sample_ids), step,
                                       L.stream_ptr()), 'ndb_philox_normal')
    _count('philox_normal')
    return out


def qsample(x0, noise, t, sqrt_abar, sqrt_1m_abar, out=None):
    L.require_device(x0, 'qsample')
    B = x0.shape[0]
    per = x0.numel() // B
    if out is None:
        out = torch.empty_like(x0)
    _begin()
    L.check(L.load().ndb_qsample(L.ptr(x0), L.ptr(noise), L.ptr(t), L.ptr(sqrt_abar), L.ptr(sqrt_1m_abar),
                                 L.ptr(out), B, per, L.stream_ptr()), 'ndb_qsample')
    _count('qsample')
    return out


def to_uint8(x_nchw, out=None, gray_invert=False):
    """[-1,1] fp32 NCHW -> uint8 NHWC exactly as the reference CLI stores it (scripts/sample.py:94,168; with
    gray_invert the one-channel double inversion of :98-100,170-171)."""
    L.require_device(x_nchw, 'to_uint8')
    x_nchw = x_nchw.contiguous()
    B, Cc, H, W = x_nchw.shape
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=x_nchw.device)
    _begin()
    L.check(L.load().ndb_to_uint8(L.ptr(x_nchw), L.ptr(out), B, Cc, H, W, 1 if gray_invert else 0, L.stream_ptr()),
            'ndb_to_uint8')
    _count('to_uint8')
    return out


def timestamp(buf, index):
    """buf: int64 device tensor; element `index` <- %globaltimer once the stream has drained (not counted in LAUNCHES)."""
    L.check(L.load().ndb_timestamp(C.c_void_p(buf.data_ptr() + 8 * index), L.stream_ptr()), 'ndb_timestamp')


def diffusion_loss(x0, x_t, noise, model_out, t, tables, var_type, loss_type, rescaled_num_steps, want_grad=True):
    """Fused training loss: -> (loss [B], grad [B,Cm,H,W] = d loss_b / d model_out or None). See ndb_diffusion_loss."""
    L.require_device(model_out, 'diffusion_loss')
    B, Cc, H, W = x0.shape
    for tns in (x0, x_t, noise, model_out):
        assert tns.is_contiguous() and tns.dtype == torch.float32
    assert t.dtype == torch.float32 and tables.dtype == torch.float32 and tables.shape[0] == L.TAB_COUNT
    p = L.LossParams()
    p.x0, p.x_t, p.noise, p.model_out = x0.data_ptr(), x_t.data_ptr(), noise.data_ptr(), model_out.data_ptr()
    p.t, p.tables = t.data_ptr(), tables.data_ptr()
    p.R, p.B, p.C, p.HW, p.Cm = tables.shape[1], B, Cc, H * W, model_out.shape[1]
    p.var_type, p.loss_type, p.rescaled_num_steps = var_type, loss_type, int(rescaled_num_steps)
    loss = torch.empty((B,), dtype=torch.float32, device=x0.device)
    grad = torch.empty_like(model_out) if want_grad else None
    p.loss = loss.data_ptr()
    p.grad_out = grad.data_ptr() if grad is not None else None
    _begin()
    L.check(L.load().ndb_diffusion_loss(C.byref(p), L.stream_ptr()), 'ndb_diffusion_loss')
    _count('diffusion_loss', 0, '', 4 * (4 * x0.numel() + model_out.numel() * (2 if want_grad else 1)))
    return loss, grad


def scale_rows_(g, upstream):
    """g[b] *= upstream[b] in place (g: [B, ...] fp32 contiguous)."""
    L.require_device(g, 'scale_rows')
    assert g.is_contiguous() and upstream.is_contiguous() and upstream.dtype == torch.float32
    _begin()
    L.check(L.load().ndb_scale_rows(L.ptr(g), L.ptr(upstream), g.shape[0], g.numel() // g.shape[0], L.stream_ptr()),
            'ndb_scale_rows')
    _count('scale_rows')
    return g
