"""The C-ABI shared library loads and exports every symbol include/b2dsp.h declares (no GPU
compute here), and the product refuses to run without a CUDA device instead of falling back."""

import ctypes
import os
import re

import numpy as np
import pytest

from dsptoolbox_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as g

        g.build()


def test_library_exports_header_symbols():
    _ensure_built()
    header = open(os.path.join(ROOT, "include", "b2dsp.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/b2dsp.h"
    assert declared == set(_native.PROTOTYPES), "ctypes prototypes and header disagree"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"


def test_host_only_entry_points():
    _ensure_built()
    lib = _native.lib()
    assert lib.b2_version() >= 100
    k, p = ctypes.c_int64(), ctypes.c_int64()
    assert lib.b2_welch_frames(2_880_000, 4096, 2048, ctypes.byref(k), ctypes.byref(p)) == 0
    assert (k.value, p.value) == (1407, 3584)
    assert lib.b2_welch_frames(48_000, 1024, 512, ctypes.byref(k), ctypes.byref(p)) == 0
    assert (k.value, p.value) == (94, 640)
    assert lib.b2_welch_frames(0, 1024, 512, None, None) == _native.E_INVAL
    assert "positive" in _native.last_error()
    assert lib.b2_welch_frame_tile(4096, _native.F64) in (1, 2, 4, 8, 16)
    assert lib.b2_welch_frame_tile(1 << 17, _native.F32) == _native.E_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _native.check(lib.b2_welch_frame_tile(1 << 17, _native.F32))
    # argument validation happens before any CUDA call
    assert lib.b2_csm_herk(None, 4, 16, 3, 0, 3, None, 0, 0, None) == _native.E_INVAL
    with pytest.raises(AssertionError):
        _native.check(lib.b2_das_prepare(None, 4, 3, 0, 1, 1, None, 0, None))
    assert lib.b2_das_operand_bytes(64, 20, _native.MAP_SIMT_F32) == 20 * 64 * 64 * 8
    # spectra layout of b2_welch_rfft: the bin -> slot map is a permutation into (wl/8 + 1) groups of 4 slots that
    # keeps bins p and wl/2 - p at slots p and wl/4 + p (what the untangling step writes together)
    for wl in (8, 16, 256, 4096, 16384):
        nc = wl // 2
        slots = [lib.b2_welch_bin_slot(wl, f) for f in range(nc + 1)]
        assert sorted(slots) == list(range(nc + 1))
        assert all(slots[p] == p and slots[nc - p] == nc // 2 + p for p in range(nc // 2))
        assert slots[nc // 2] == nc
        assert lib.b2_welch_spectra_elems(wl, 7, 48) == 3 * (wl // 8 + 1) * 7 * 64
    assert lib.b2_welch_bin_slot(4096, 2049) == -1 and lib.b2_welch_bin_slot(4000, 0) == -1
    assert lib.b2_welch_spectra_elems(4096, 7, 40) == -1
    assert lib.b2_heig_max_order() == 80
    assert lib.b2_heig_jacobi(None, 4, 1, None, None, None, None) == _native.E_INVAL
    assert lib.b2_cleansc_component(None, None, 4, 1, None, None, None, None, None, None, None, 0, 0.5, 0, None, None, None) == _native.E_INVAL
    assert lib.b2_das_rank1_update(None, None, 4, 1, None, 343.0, None, None, None, 10, 0, 0, -0.5, None, None) == _native.E_INVAL


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dsptoolbox_b200 as dsp

    s = dsp.Signal(None, np.random.default_rng(0).standard_normal((2000, 3)), 8000)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.get_csm()
