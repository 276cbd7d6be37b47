"""vviewer_b200 — B200-native drop-in for vengine's offline GPU path-tracing render path.

The product is native: ``_lib/libptc_cuda.so`` (hand-written sm_100a CUDA behind ``include/ptc.h``) and
``_lib/libvengine_host.so`` (the C++ vengine-shaped host library).  This Python package only binds them
for tests and bench.py.
"""
from . import capi  # noqa: F401
