"""The C-ABI libraries load and export every symbol the headers declare (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"PTC_API[^;]*?\b(%s_\w+)\s*\(" % prefix, text)))


def test_header_symbols_listed(capi):
    assert declared("ptc.h", "ptc") == sorted(capi.PTC_SYMBOLS)
    assert declared("vengine_host.h", "vh") == sorted(capi.VH_SYMBOLS)


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_ptc_library_exports(capi, which):
    path = capi.CUDA_LIB if which == "cuda" else capi.ORACLE_LIB
    lib = ctypes.CDLL(path)
    for sym in capi.PTC_SYMBOLS:
        assert hasattr(lib, sym), "%s does not export %s" % (path, sym)
    capi._declare_ptc(lib)
    name = lib.ptc_backend_name().decode()
    assert name == ("cuda-sm_100a" if which == "cuda" else "cpu-oracle")


def test_host_library_exports(capi):
    lib = ctypes.CDLL(capi.HOST_LIB)
    for sym in capi.VH_SYMBOLS:
        assert hasattr(lib, sym)


def test_pod_sizes(capi):
    # sizes the reference's UBO / SSBO records have
    assert ctypes.sizeof(capi.ptc_vertex) == 68
    assert ctypes.sizeof(capi.ptc_instance) == 128
    assert ctypes.sizeof(capi.ptc_material) == 128
    assert ctypes.sizeof(capi.ptc_light_data) == 64
    assert ctypes.sizeof(capi.ptc_light_instance) == 64
    assert ctypes.sizeof(capi.ptc_scene_data) == 304


def test_cuda_library_fails_loudly_without_gpu(capi):
    """There is no CPU fallback: on a box without a GPU ptc_create must fail with a message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load_cuda()
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        capi.Context(lib)


def test_product_does_not_reference_oracle():
    """The product sources (package + include) must never include, link or load anything under oracle/."""
    bad = []
    for base in ("vviewer_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_lib" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    s = line.strip()
                    if ("#include" in s and "oracle" in s) or "dlopen(\"oracle" in s:
                        bad.append((f, s))
    assert not bad, bad
