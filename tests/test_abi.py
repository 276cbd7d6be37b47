"""The C-ABI libraries load and export every symbol the headers declare (no compute without a GPU)."""
import ctypes
import os
import re

import pytest
from oracle import loader as oracle_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"PTC_API[^;]*?\b(%s_\w+)\s*\(" % prefix, text)))


def test_header_symbols_listed(capi):
    assert declared("ptc.h", "ptc") == sorted(capi.PTC_SYMBOLS)
    assert declared("vengine_host.h", "vh") == sorted(capi.VH_SYMBOLS)


@pytest.mark.parametrize("which", ["cuda", "oracle"])
def test_ptc_library_exports(capi, which):
    path = capi.CUDA_LIB if which == "cuda" else oracle_loader.ORACLE_LIB
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
    """The product (package, headers, its binaries' sources) must never include, link, name or load anything under oracle/,
    and must not be steerable to another backend: no path string, no ctypes / dlopen of a caller-named library, no --backend."""
    bad = []
    word = re.compile(r"oracle", re.I)
    for base in ("vviewer_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_lib" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if not f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h", ".py")):
                    continue
                path = os.path.join(dirpath, f)
                for ln, line in enumerate(open(path, errors="ignore").read().splitlines(), 1):
                    s = line.strip()
                    code = s.split("//")[0]
                    in_comment = s.startswith(("/*", "*", "//", "#", '"""')) or "/*" in s and s.index("/*") < (s.lower().index("oracle") if "oracle" in s.lower() else 0)
                    # 1. the word may appear in prose (comments / docstrings explaining what the tests check), never in code or strings
                    if word.search(code) and not in_comment and not (f.endswith(".py") and (s.startswith(("#", '"', "'")) or '"""' in s)):
                        if re.search(r"[\"'][^\"']*oracle[^\"']*[\"']", s, re.I) or "#include" in s or "import" in s or "load" in s.lower():
                            bad.append((path, ln, s))
                    # 2. no run-time loading of a library the caller names
                    if "dlopen(" in code and not any(k in code for k in ("libPath.c_str()", "(n, RTLD")):
                        bad.append((path, ln, s))
                    if "--backend" in s and "unknown" not in s:
                        bad.append((path, ln, s))
                    if f.endswith(".py") and re.search(r"CDLL\(|cdll\.LoadLibrary", code) and "path" not in code and "HOST_LIB" not in code:
                        bad.append((path, ln, s))
    assert not bad, bad
    # the two dlopen sites that exist open FIXED names: the CUDA core next to the host library, and NCCL
    host = open(os.path.join(ROOT, "vviewer_b200", "host", "vengine.cpp")).read()
    assert 'm_backend.load(selfDir() + "/libptc_cuda.so"' in host and host.count("dlopen(") == 1
    core = open(os.path.join(ROOT, "vviewer_b200", "csrc", "ptc_cuda.cu")).read()
    assert core.count("dlopen(") == 1 and '"libnccl.so.2"' in core
    # the Python package exposes no oracle loader and the engine takes no backend argument
    from vviewer_b200 import capi
    assert not hasattr(capi, "ORACLE_LIB") and not hasattr(capi, "load_oracle")
    import inspect
    assert "backend" not in inspect.signature(capi.HostEngine.__init__).parameters
    assert "backend_lib" not in open(os.path.join(ROOT, "include", "vengine_host.h")).read()


@pytest.mark.parametrize("header", ["ptc.h", "vengine_host.h"])
def test_public_headers_are_plain_c(header, tmp_path):
    """the drop-in boundary is a C ABI: the headers must compile as C11 (what a cgo / JNI / ctypes-generator binding consumes) and as
    C++17 (the reference's language) without any other include"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include "%s"\nint main(void) { return 0; }\n' % header)
    for cc, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        cmd = [cc, std, "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(root, "include")]
        if cc == "g++":
            cmd += ["-x", "c++"]
        r = subprocess.run(cmd + [str(src)], capture_output=True, text=True)
        assert r.returncode == 0, "%s %s: %s" % (cc, header, r.stderr[:2000])
