"""Loader of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module
(tests/test_abi.py::test_product_does_not_reference_oracle guards the product tree).  The product
(vviewer_b200/, include/) never loads, names or links anything under oracle/.
"""
import os

from vviewer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

_oracle = None


def load_oracle():
    """CPU restatement of the reference behind the same C-ABI (include/ptc.h) as the product."""
    global _oracle
    if _oracle is None:
        _oracle = capi.load_ptc(ORACLE_LIB)
    return _oracle


def oracle_context(eng, hierarchy=None):
    """A ptc context of the ORACLE holding the scene the host engine `eng` currently describes (uploaded, accel built)."""
    ctx = capi.Context(load_oracle())
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel(hierarchy)
    return ctx


def oracle_render(eng, want_aovs=True, with_stats=False):
    """RendererPathTracing::render() of the host engine's scene, computed by the oracle instead of the product:
    same flattened scene, same render parameters, through the raw C-ABI."""
    ctx = oracle_context(eng)
    try:
        out = ctx.render(eng.render_params(), want_aovs=want_aovs)
        st = ctx.stats()
    finally:
        ctx.close()
    return (out, st) if with_stats else out
