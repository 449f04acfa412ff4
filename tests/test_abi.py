"""CPU: the C-ABI library is built in-tree, loads, and exports every symbol include/msckf_b200.h declares.
No compute calls are made here (no GPU)."""
import ctypes
import os
import re

import pytest

from monocular_visual_inertial_msckf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "msckf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msckf_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports():
    if not os.path.exists(_lib.lib_path()):
        pytest.fail("libmsckf_b200.so not built: run `make` at the repo root")
    lib = ctypes.CDLL(_lib.lib_path())
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.msckf_abi_version() == _lib.ABI_VERSION == 10
    assert lib.msckf_max_obs_per_feature() >= 61
    assert lib.msckf_max_cameras(0) >= 31 and lib.msckf_max_cameras(1) >= 61


def test_argument_checks_without_gpu():
    lib = _lib.load_library()
    d = _lib.Dims(4, 8, 8, 4, 15 + 24)
    assert lib.msckf_update_workspace_bytes(ctypes.byref(d), 0) > 0
    bad = _lib.Dims(4, 8, 8, 4, 10)                      # ldp too small
    assert lib.msckf_update_workspace_bytes(ctypes.byref(bad), 0) == 0
    p = _lib.Params()
    rc = lib.msckf_update_batched(ctypes.byref(d), ctypes.byref(p), None, 0, *([None] * 19), None, None, None, None,
                                  0, None)
    assert rc == -1                                       # MSCKF_EINVAL on null pointers, before any launch


def test_no_cpu_fallback():
    import torch
    from monocular_visual_inertial_msckf_b200 import BatchedMSCKF, UpdateParams
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        BatchedMSCKF(1, 1, 2, 2, UpdateParams())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "monocular_visual_inertial_msckf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"
