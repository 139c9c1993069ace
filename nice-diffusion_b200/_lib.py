"""ctypes binding of libndb200.so (the C-ABI declared in include/ndb200.h).

Thin by design: raw device pointers (``tensor.data_ptr()``), sizes and the current CUDA stream go in, an
int status comes out and is turned into a Python exception. Nothing here computes anything, and there is
no fallback: if the library is missing the import error says how to build it.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NDB_LIB: another build of the same library (A/B timing of two builds on one GPU box, bring-up builds)
LIB_PATH = os.environ.get('NDB_LIB') or os.path.join(_HERE, 'libndb200.so')

MAX_SEGMENTS = 9
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3
RES_NONE, RES_SAME, RES_UP2, RES_DOWN2 = 0, 1, 2, 3
OUT_F32_NHWC, OUT_F16_NHWC, OUT_F32_NCHW = 0, 1, 2
RESAMPLE_NONE, RESAMPLE_UP2, RESAMPLE_DOWN2 = 0, 1, 2
VAR_FIXED, VAR_LEARNED, VAR_LEARNED_INTERP = 0, 1, 2
(TAB_SQRT_RECIP_ABAR, TAB_SQRT_RECIPM1_ABAR, TAB_POST_COEF_X0, TAB_POST_COEF_XT, TAB_LOGVAR_MIN,
 TAB_LOGVAR_MAX, TAB_ABAR, TAB_ABAR_PREV, TAB_SQRT_1M_ABAR, TAB_LOG_POST_VAR_CLIPPED, TAB_COUNT) = range(11)
LOSS_SIMPLE, LOSS_KL, LOSS_KL_RESCALED, LOSS_HYBRID = 0, 1, 2, 3

FMT_FP16, FMT_BF16 = 0, 1
# fp32-class mode of the model (DiffusionModel.operand_format): every tensor-core operand is an fp16 (hi, lo) pair and every
# product is evaluated as hi*hi + lo*hi + hi*lo; the kernels themselves run in FMT_FP16
FMT_FP32X = 2
SPLIT_ACC_CHUNK = 2           # K blocks per tensor-core accumulation chunk in the fp32-class mode (8 MMA steps)
SPLIT_WEIGHT_SCALE = 256.0   # weights are packed times 2^8 so that their low halves stay normal fp16 numbers


class ConvSegment(C.Structure):
    _fields_ = [('act', C.c_void_p), ('weight', C.c_void_p), ('C', C.c_int32), ('taps', C.c_int32)]


class ConvParams(C.Structure):
    _fields_ = [('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('cout', C.c_int32), ('cout_pad', C.c_int32), ('num_segments', C.c_int32),
                ('seg', ConvSegment * MAX_SEGMENTS),
                ('bias', C.c_void_p), ('residual', C.c_void_p),
                ('residual_mode', C.c_int32), ('out_kind', C.c_int32),
                ('out', C.c_void_p),
                ('operand_format', C.c_int32), ('block_n', C.c_int32),
                ('stat_parts', C.c_void_p), ('stat_smax', C.c_int32), ('out_scale', C.c_float), ('acc_chunk_kblocks', C.c_int32)]


class GnApplyParams(C.Structure):
    _fields_ = [('x0', C.c_void_p), ('C0', C.c_int32),
                ('x1', C.c_void_p), ('C1', C.c_int32),
                ('N', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('parts0', C.c_void_p), ('counts0', C.c_void_p), ('smax0', C.c_int32),
                ('parts1', C.c_void_p), ('counts1', C.c_void_p), ('smax1', C.c_int32),
                ('gamma', C.c_void_p), ('beta', C.c_void_p),
                ('scale', C.c_void_p), ('shift', C.c_void_p), ('ld_ss', C.c_int32),
                ('pre_add', C.c_void_p), ('ld_pre', C.c_int32),
                ('silu', C.c_int32), ('resample', C.c_int32),
                ('out', C.c_void_p), ('raw16', C.c_void_p), ('x0_is16', C.c_int32), ('pool32', C.c_void_p),
                ('operand_format', C.c_int32), ('eps', C.c_float), ('out_lo', C.c_void_p), ('raw_lo', C.c_void_p),
                ('cpg', C.c_int32)]


class StepParams(C.Structure):
    _fields_ = [('x_t', C.c_void_p), ('model_out', C.c_void_p), ('null_out', C.c_void_p),
                ('t', C.c_void_p), ('tables', C.c_void_p),
                ('R', C.c_int32), ('B', C.c_int32), ('C', C.c_int32), ('HW', C.c_int32), ('Cm', C.c_int32),
                ('var_type', C.c_int32), ('ddim', C.c_int32),
                ('eta', C.c_double), ('cfg_strength', C.c_double),
                ('grad', C.c_void_p), ('clf_strength', C.c_double),
                ('clip_x0', C.c_int32),
                ('noise', C.c_void_p), ('seed', C.c_uint64), ('sample_ids', C.c_void_p),
                ('x_prev', C.c_void_p), ('x_prev2', C.c_void_p), ('pred_x0', C.c_void_p)]


class LossParams(C.Structure):
    _fields_ = [('x0', C.c_void_p), ('x_t', C.c_void_p), ('noise', C.c_void_p), ('model_out', C.c_void_p),
                ('t', C.c_void_p), ('tables', C.c_void_p),
                ('R', C.c_int32), ('B', C.c_int32), ('C', C.c_int32), ('HW', C.c_int32), ('Cm', C.c_int32),
                ('var_type', C.c_int32), ('loss_type', C.c_int32), ('rescaled_num_steps', C.c_int32),
                ('loss', C.c_void_p), ('grad_out', C.c_void_p)]


# name -> (restype, argtypes); must list every symbol include/ndb200.h declares (tests check this)
_I, _P, _I64, _U64 = C.c_int, C.c_void_p, C.c_int64, C.c_uint64
SIGNATURES = {
    'ndb_version': (_I, []),
    'ndb_device_supported': (_I, []),
    'ndb_last_error': (C.c_char_p, []),
    'ndb_conv_gemm': (_I, [C.POINTER(ConvParams), _P]),
    'ndb_conv_stats_layout': (_I, [_I, _I, _I, _I, _I, _I, _I, _P]),
    'ndb_conv_stats_supported': (_I, [_I, _I, _I]),
    'ndb_pack_conv_weight': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, C.c_float, _I, _P]),
    'ndb_split16': (_I, [_P, _P, _P, _I64, _I, _P]),
    'ndb_gn_stats': (_I, [_P, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    'ndb_gn_stats_workspace_bytes': (_I64, [_I, _I, _I]),
    'ndb_gn_apply': (_I, [C.POINTER(GnApplyParams), _P]),
    'ndb_attention': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    'ndb_attention_split': (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_embed': (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    'ndb_stem_conv': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_resample_f32': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_cast_nhwc16': (_I, [_P, _P, _I64, _I, _P]),
    'ndb_im2col_s2': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_stem_im2col': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_sampler_step': (_I, [C.POINTER(StepParams), _P]),
    'ndb_advance_step': (_I, [_P, _I, _P, _I, _P, _P]),
    'ndb_philox_normal': (_I, [_P, _I, _I64, _U64, _P, C.c_int32, _P]),
    'ndb_qsample': (_I, [_P, _P, _P, _P, _P, _P, _I, _I64, _P]),
    'ndb_to_uint8': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'ndb_timestamp': (_I, [_P, _P]),
    'ndb_diffusion_loss': (_I, [C.POINTER(LossParams), _P]),
    'ndb_scale_rows': (_I, [_P, _P, _I64, _I64, _P]),
}

_lib = None


class NdbError(RuntimeError):
    pass


def load():
    """dlopen libndb200.so (once) and attach prototypes. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libndb200.so not found at {} -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or nice-diffusion_b200/csrc/build.sh (needs nvcc; there is no CPU fallback)'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != OK:
        names = {ERR_INVALID: 'invalid argument', ERR_CUDA: 'CUDA error', ERR_UNSUPPORTED: 'unsupported'}
        raise NdbError('{} failed: {} ({})'.format(what, names.get(status, 'error'), status))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def require_device(t, what):
    if not t.is_cuda:
        raise NdbError('{}: tensor is on {} -- the ndb200 path runs on a B200 only (no CPU fallback)'.format(
            what, t.device))
    # kernels are enqueued on the CURRENT device's current stream (stream_ptr): a tensor of another GPU would be
    # launched on the wrong device. One process per GPU (torch.cuda.set_device) is the supported model.
    if t.device.index != torch.cuda.current_device():
        raise NdbError('{}: tensor is on cuda:{} but the current device is cuda:{} -- call torch.cuda.set_device() '
                       '(one process per GPU)'.format(what, t.device.index, torch.cuda.current_device()))
