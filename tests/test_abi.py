"""No-GPU checks of the drop-in boundary: libqsmrt.so loads, exports every
symbol include/qsmrt.h declares, and fails loudly (no CPU fallback) when
there is no CUDA device."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "qsmrt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qsmrt_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    from pyqsm_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/qsmrt.h but not exported"
    assert set(declared) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.qsmrt_abi_version() == _lib.ABI_VERSION == 2


def test_no_torch_in_abi():
    """The boundary is plain C: the .so must not link libtorch / libc10."""
    import subprocess
    from pyqsm_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from pyqsm_b200 import _lib, RaycastingScene
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.qsmrt_scene_create(0, C.byref(h)) != 0
    assert b"no CPU fallback" in lib.qsmrt_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RaycastingScene()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under pyqsm_b200/ or tools/ may reference it."""
    for top in ("pyqsm_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, fn)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                    assert "libqsmrt_oracle" not in text, fn
