"""ctypes binding of libci_b200.so (the C-ABI in include/ci_b200.h).

There is NO CPU fallback: if the shared object is missing, cannot be loaded, or CUDA is not
available, every entry point raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libci_b200.so')
_lib = None


class NativeError(RuntimeError):
    pass


class CiLayout(ctypes.Structure):
    # S2 / r2: second Seasonal component of a custom model (0 = none); see include/ci_b200.h
    _fields_ = [(n, ctypes.c_int32) for n in ('N', 'G', 'T', 'Tf', 'k', 'S', 'r', 'Tpad', 'flags', 'S2', 'r2')]


class CiData(ctypes.Structure):
    _fields_ = [('d_y', ctypes.c_void_p), ('d_X', ctypes.c_void_p), ('d_sconst', ctypes.c_void_p),
                ('d_XT', ctypes.c_void_p)]


def lib():
    """Load the shared object (building is `python -m tfcausalimpact_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(
                f'{_LIB_PATH} not found: build it with `python -m tfcausalimpact_b200.build` '
                '(nvcc, sm_100a). There is no CPU fallback.')
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ci_last_error_string.restype = ctypes.c_char_p
        for name in ('ci_eval_workspace_bytes', 'ci_vi_workspace_bytes', 'ci_hmc_workspace_bytes',
                     'ci_predict_workspace_bytes', 'ci_reduce_workspace_bytes', 'ci_fstate_floats', 'ci_decompose_workspace_bytes'):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = ctypes.c_size_t
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError('tfcausalimpact_b200 needs a CUDA device (sm_100a); there is no CPU fallback.')


def check(rc, what=''):
    if rc != 0:
        msg = lib().ci_last_error_string().decode(errors='replace')
        raise NativeError(f'{what} failed with code {rc}: {msg}')


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), 'device pointer arguments must be contiguous CUDA tensors'
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


FLAG_DENSE_REG, FLAG_LLT, FLAG_OBS_LOGNORMAL, FLAG_LEVEL_LOGNORMAL = 1, 2, 4, 8


def make_layout(N, G, T, Tf, k, S=1, r=1, Tpad=None, flags=0, S2=0, r2=0):
    if Tpad is None:
        Tpad = (T + Tf + 3) // 4 * 4
    return CiLayout(N, G, T, Tf, k, S, r, Tpad, flags, S2, r2)


def num_params(k, S, flags=0, S2=0):
    trend = 2 if flags & FLAG_LLT else 1
    reg = 0 if k == 0 else (k if flags & FLAG_DENSE_REG else 2 + 3 * k)
    return 1 + trend + reg + (1 if S > 1 else 0) + (1 if S2 > 1 else 0)
