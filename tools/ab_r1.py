"""A/B of the current library against the round-1 library (built from git 25bcd3b into _lib/libptc_cuda_r1.so) on the SAME box,
alternating, through the part of the C-ABI both share.  usage: python tools/ab_r1.py Scene[:batches] ...
AB_LIBS=label=path,label=path compares other builds of the library instead (experiment variants under _lib/); a path may carry
@KEY=VALUE to run that entry with an environment variable set (run-time switches of one build)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vviewer_b200 import capi  # noqa: E402


class OldStats(C.Structure):  # ptc_stats of round 1
    _fields_ = capi.ptc_stats._fields_[:17]


def run(path, desc, rp, batches):
    path, _, setting = path.partition("@")
    if setting:
        os.environ[setting.split("=", 1)[0]] = setting.split("=", 1)[1]
    try:
        return run_lib(path, desc, rp, batches)
    finally:
        if setting:
            os.environ.pop(setting.split("=", 1)[0], None)


def run_lib(path, desc, rp, batches):
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp = C.c_void_p
    lib.ptc_create.argtypes = [C.POINTER(vp), vp, C.c_int]
    lib.ptc_upload_scene.argtypes = [vp, C.POINTER(capi.ptc_scene_desc)]
    lib.ptc_build_accel.argtypes = [vp]
    lib.ptc_render.argtypes = [vp, C.POINTER(capi.ptc_render_params), vp, vp, vp]
    lib.ptc_get_stats.argtypes = [vp, vp]
    lib.ptc_destroy.argtypes = [vp]
    lib.ptc_last_error.argtypes = [vp]
    lib.ptc_last_error.restype = C.c_char_p
    ctx = vp()
    assert lib.ptc_create(C.byref(ctx), None, 0) == 0
    assert lib.ptc_upload_scene(ctx, desc) == 0, lib.ptc_last_error(ctx)
    assert lib.ptc_build_accel(ctx) == 0
    rad = np.zeros(rp.width * rp.height * 4, np.float32)
    out = []
    buf = (C.c_uint8 * 512)()
    for flags in (0, 0, 0, capi.PTC_FLAG_TIME_KERNELS):
        rp.samples = batches * rp.batch_size
        rp.flags = flags
        assert lib.ptc_render(ctx, C.byref(rp), rad.ctypes.data_as(vp), None, None) == 0, lib.ptc_last_error(ctx)
        lib.ptc_get_stats(ctx, buf)
        st = OldStats.from_buffer_copy(bytes(buf)[:C.sizeof(OldStats)])
        out.append((st.segments / st.render_ms / 1e3, st.render_ms / batches, st.trace_ms / batches, st.shade_ms / batches, st.shadow_ms / batches))
    lib.ptc_destroy(ctx)
    return out, float(rad.reshape(-1, 4)[:, :3].mean())


LIBS = [("r1", os.path.join(capi.LIB_DIR, "libptc_cuda_r1.so")), ("now", capi.CUDA_LIB)]
if os.environ.get("AB_LIBS"):
    LIBS = [(kv.split("=", 1)[0], os.path.join(ROOT, kv.split("=", 1)[1])) for kv in os.environ["AB_LIBS"].split(",")]

for spec in sys.argv[1:] or ["Cornell:8", "Atrium:4"]:
    name, _, b = spec.partition(":")
    batches = int(b) if b else 4
    eng = capi.HostEngine()
    eng.build_scene(name, texture_size=1024 if name in ("Atrium", "Fog") else 512)
    desc = eng.scene_desc()
    for rnd in range(2):
        for label, path in LIBS:
            res, mean = run(path, desc, eng.render_params(), batches)
            best = max(r[0] for r in res[:3])
            t = res[3]
            print("%-12s %-5s best of 3: %8.1f Mseg/s (%.2f ms/batch) | timed: extend %.2f shade %.2f chains %.2f | mean %.6f" % (
                name, label, best, min(r[1] for r in res[:3]), t[2], t[3], t[4], mean), flush=True)
    eng.close()
