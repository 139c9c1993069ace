"""B200-native sampling hot path for edogariu/nice-diffusion.

Host-side mirror of the reference's Python API (``model.DiffusionModel``, ``diffusion.Diffusion``,
``utils.make_argparser`` / ``get_dicts_from_args``, ``default_args``) on top of ``libndb200.so`` -- a C-ABI
library of hand-written sm_100a CUDA kernels (``csrc/``, declarations in ``include/ndb200.h``).
There is no CPU or PyTorch-eager fallback: model forwards require a B200 and the built library.
"""
__version__ = '0.1.0'
