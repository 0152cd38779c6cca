"""ctypes binding of libb2dsp.so (C ABI declared in include/b2dsp.h).

There is NO CPU fallback: if the library is missing or a call fails, an exception is raised.
ctypes releases the GIL for the duration of every call; the calls only enqueue kernels on
the given CUDA stream.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2DSP_LIB lets a developer A/B-test an alternative build of the same library (never a fallback)
LIB_PATH = os.environ.get("B2DSP_LIB") or os.path.join(_HERE, "csrc", "libb2dsp.so")

# precision / mode codes (include/b2dsp.h)
F64, F32 = 0, 1
MAP_SIMT_F64, MAP_SIMT_F32, MAP_TC_TF32, MAP_TC_TF32X3, MAP_TC_MIXED, MAP_TC_F16X3 = 0, 1, 2, 3, 4, 5
STEER_CLASSIC, STEER_INVERSE, STEER_TRUE_POWER, STEER_TRUE_LOCATION = 0, 1, 2, 3
E_INVAL, E_UNSUPPORTED, E_CUDA = -1, -2, -3

# every exported symbol with (restype, argtypes); tests check that the library exports all
PROTOTYPES = {
    "b2_version": (c_int, []),
    "b2_last_error": (c_char_p, []),
    "b2_welch_frames": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "b2_welch_frame_tile": (c_int, [c_int, c_int]),
    "b2_welch_spectra_elems": (c_int64, [c_int, c_int, c_int64]),
    "b2_welch_bin_slot": (c_int, [c_int, c_int]),
    "b2_fft_twiddles": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "b2_prepare_channels": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p,
                                    c_int64, c_int, c_void_p]),
    "b2_welch_rfft": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_int,
                              c_double, c_void_p, c_int64, c_int, c_void_p]),
    "b2_csm_herk": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b2_csm_finalize": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_void_p]),
    "b2_welch_autopower": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_int, c_void_p]),
    "b2_auto_finalize": (c_int, [c_void_p, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p, c_void_p]),
    "b2_cleansc_component": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     POINTER(c_double), c_void_p, c_int, c_double, c_int, c_void_p, c_void_p, c_void_p]),
    "b2_das_rank1_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_double, c_void_p, POINTER(c_double),
                                    c_void_p, c_int64, c_int, c_int, c_double, c_void_p, c_void_p]),
    "b2_heig_max_order": (c_int, []),
    "b2_heig_jacobi": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2_fft_plan_bytes": (c_int64, [c_int64]),
    "b2_fft_work_bytes": (c_int64, [c_int64, c_int]),
    "b2_fft_plan": (c_int, [c_int64, c_void_p, c_void_p, c_void_p]),
    "b2_rfft_any": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int, c_int64, c_int64, c_int, c_void_p, c_int,
                            c_void_p, c_void_p, c_double, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "b2_delay_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "b2_csm_rank1": (c_int, [c_void_p, c_int, c_int64, c_double, c_int, c_int, c_void_p, c_void_p]),
    "b2_das_operand_bytes": (c_int64, [c_int, c_int, c_int]),
    "b2_das_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "b2_das_map": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_double, c_void_p, POINTER(c_double),
                           c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "b2_steering_vector": (c_int, [c_void_p, c_int, c_void_p, c_int, POINTER(c_double), c_void_p, c_int64,
                                   c_int, c_void_p, c_void_p]),
}


class NativeError(RuntimeError):
    """A b2dsp call failed (CUDA error or internal failure)."""


class NativeLibraryMissing(ImportError):
    pass


_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or dsptoolbox_b200/csrc/build.sh). The B200 path has no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().b2_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "") -> int:
    """Translate a status code into the exception the reference would raise for that input.

    Argument validation in the reference is done with bare `assert`, so
    B2_EINVAL maps to AssertionError; unsupported-but-valid requests to NotImplementedError.
    """
    if status >= 0:
        return status
    msg = f"{what}: {last_error()}" if what else last_error()
    if status == E_INVAL:
        raise AssertionError(msg)
    if status == E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NativeError(msg)


def ptr(t) -> int:
    """Device pointer of a torch tensor (or None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_handle(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def vec3(xyz) -> ctypes.Array:
    return (c_double * 3)(float(xyz[0]), float(xyz[1]), float(xyz[2]))
