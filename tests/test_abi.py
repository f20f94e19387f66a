"""The C-ABI library loads and exports every symbol include/ykgpu.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from yunikorn_k8shim_b200 import build
    return ctypes.CDLL(build.build())


def declared_functions(header="ykgpu.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(yk_[a-z_]+)\s*\(", src))
    names.discard("yk_allgather_fn")
    return sorted(names)


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("yk_create", "yk_destroy", "yk_nodes_upsert", "yk_asks_upsert", "yk_cycle", "yk_release", "yk_evaluate"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in ykgpu.h but not exported by libykgpu.so"


def test_python_stub_lists_the_same_symbols():
    from yunikorn_k8shim_b200 import EXPORTS
    from yunikorn_k8shim_b200.dictionary import DICT_EXPORTS
    assert sorted(EXPORTS) == declared_functions()
    assert sorted(DICT_EXPORTS) == declared_functions("ykgpu_dict.h")
    from yunikorn_k8shim_b200.podres import POD_EXPORTS
    assert sorted(POD_EXPORTS) == declared_functions("ykgpu_pod.h")


def test_pod_symbols_are_exported(lib):
    for name in declared_functions("ykgpu_pod.h"):
        assert hasattr(lib, name), f"{name} declared in ykgpu_pod.h but not exported by libykgpu.so"


def test_dictionary_symbols_are_exported(lib):
    for name in declared_functions("ykgpu_dict.h"):
        assert hasattr(lib, name), f"{name} declared in ykgpu_dict.h but not exported by libykgpu.so"


def test_abi_version_and_strerror(lib):
    lib.yk_abi_version.restype = ctypes.c_uint32
    assert lib.yk_abi_version() == 3
    lib.yk_strerror.restype = ctypes.c_char_p
    assert b"CUDA" in lib.yk_strerror(-2)
    assert lib.yk_strerror(0) == b"ok"


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product must fail loudly, not fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from yunikorn_k8shim_b200 import Engine, YkError
    with pytest.raises(YkError) as ei:
        Engine()
    assert ei.value.code == -2


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may mention it."""
    pkg = os.path.join(ROOT, "yunikorn_k8shim_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_ctypes" not in txt and "yk_oracle" not in txt and "py_oracle" not in txt, f


def test_python_stub_array_helpers():
    """engine.py's argument marshalling (no library call): row-major [n][D] and column-major [D][n] inputs both become the
    ABI's contiguous [D][n] int64; wrong shapes are rejected; typed arrays pass through without a copy"""
    import numpy as np
    from yunikorn_k8shim_b200 import engine as E
    a = np.arange(12, dtype=np.int64).reshape(4, 3)            # [n=4][D=3]
    c = E._colmajor(a, 3, 4)
    assert c.shape == (3, 4) and c.flags["C_CONTIGUOUS"] and np.array_equal(c, a.T)
    c2 = E._colmajor(np.ascontiguousarray(a.T), 3, 4)
    assert np.array_equal(c2, a.T)
    with pytest.raises(ValueError):
        E._colmajor(np.zeros((5, 3), dtype=np.int64), 3, 4)
    x = np.arange(5, dtype=np.uint32)
    assert E._arr(x, np.uint32, 5) is x                         # no copy when the dtype already matches
    assert E._arr([1, 2, 3], np.int32).dtype == np.int32
    with pytest.raises(ValueError):
        E._arr(x, np.uint32, 4)
    assert E._arr(None, np.uint32) is None and E._p(None) is None
