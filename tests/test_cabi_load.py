"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU, and
exports every symbol include/antsrl_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from antsrl_b200 import _cabi
    return _cabi.load_library()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "antsrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ants_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from antsrl_b200 import _cabi
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared


def test_abi_version_and_struct_sizes(lib):
    from antsrl_b200 import _cabi
    assert lib.ants_abi_version() == _cabi.ABI_VERSION
    # layout of the ctypes mirrors must match the C structs (computed by hand from the header)
    assert ctypes.sizeof(_cabi.AntsHostState) == 23 * 8 + 8 + 4 + 4
    assert ctypes.sizeof(_cabi.AntsStats) == 10 * 8
    assert ctypes.sizeof(_cabi.AntsConfig) % 8 == 0
    assert ctypes.sizeof(_cabi.AntsPackedLayout) == 5 * 4 + 4 + 8 + 8 * 4 + 2 * 4 + 4 + 228


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from antsrl_b200 import AntsError, BatchedAnts, make_config
    with pytest.raises(AntsError):
        BatchedAnts(make_config(32, 32, 4), 1)
    # and straight through the C ABI
    from antsrl_b200.batch import build_c_config
    cfg = build_c_config(make_config(32, 32, 4), 1)
    h = ctypes.c_void_p()
    rc = lib.ants_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.ants_last_error() or b"device" in lib.ants_last_error()


def test_sass_is_sm100a():
    import subprocess
    from antsrl_b200 import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
