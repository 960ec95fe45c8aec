"""The C-ABI library loads without a GPU and exports every symbol include/atacom_b200.h declares;
argument validation happens before any CUDA call."""
import ctypes
import os
import re

import pytest

from rl_on_manifold_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "atacom_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(atacom_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), "missing export: " + name
        assert name in _lib.SIGNATURES, "no ctypes prototype for " + name
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_error_strings():
    assert "sm_100a" in _lib.version()
    assert _lib.lib.atacom_error_string(0) == b"ok"
    assert b"null" in _lib.lib.atacom_error_string(-1)


def test_params_struct_size_matches_header():
    # 64 gains + dt, tol + 4 ints + 24 doubles
    assert ctypes.sizeof(_lib.AtacomParams) == 4 * 66 + 4 * 4 + 8 * 24


def test_default_params_iiwa():
    p = _lib.default_params("iiwa", 6)
    assert abs(p.K_f[0] - 0.1) < 1e-7 and p.K_c[0] == 240.0 and abs(p.rref_tol - 0.05) < 1e-8
    assert p.env[3] == 0.1505 and p.env[0] == -1.51
    with pytest.raises(_lib.AtacomError):
        _lib.default_params("iiwa", 5)


def test_argument_validation_without_gpu():
    p = _lib.default_params("circle")
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(16)
    # B = 0 with valid pointers is a no-op; null pointers and bad params are rejected before any launch
    assert _lib.lib.atacom_circle_step(one, one, one, one, one, one, null, null, 0, ctypes.byref(p), null) == 0
    assert _lib.lib.atacom_circle_step(null, one, one, one, one, one, null, null, 4, ctypes.byref(p), null) == -1
    assert _lib.lib.atacom_circle_step(one, one, one, one, one, one, null, null, 4, None, null) == -1
    assert _lib.lib.atacom_circle_step(null, null, null, null, null, null, null, null, 0, ctypes.byref(p), null) == 0
    bad = p.copy()
    bad.variant = 7
    assert _lib.lib.atacom_circle_step(one, one, one, one, one, one, null, null, 4, ctypes.byref(bad), null) == -3
    assert _lib.lib.atacom_iiwa_step(5, one, one, one, one, one, one, null, null, 4, ctypes.byref(p), null) == -2
    assert _lib.lib.atacom_point_reach_step(5, one, one, one, one, one, one, one, one, null, null, 4,
                                            ctypes.byref(p), null) == -2
    assert _lib.lib.atacom_generic_supported(6, 1, 11) == 1 and _lib.lib.atacom_generic_supported(5, 3, 1) == 0


def test_cpu_tensors_are_rejected():
    import torch
    from rl_on_manifold_b200 import projection
    p = _lib.default_params("circle")
    z = torch.zeros(4, 2)
    with pytest.raises(ValueError, match="CUDA"):
        projection.step("circle", z, z, torch.zeros(4, 1), torch.zeros(4, 1), p)
