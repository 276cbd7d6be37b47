"""Acceleration-structure build time per scene (CUDA events inside ptc_build_accel), first build (allocations) and rebuilds.
usage: python tools/build_time.py Scene[:scale] ...   (BUILD_LIB=path times another build of the library, e.g. an experiment variant)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi  # noqa: E402

cuda = capi.load_ptc(os.environ["BUILD_LIB"]) if os.environ.get("BUILD_LIB") else capi.load_cuda()
for spec in sys.argv[1:] or ["Cornell", "Atrium", "Instanced:0.25"]:
    name, _, scale = spec.partition(":")
    eng = capi.HostEngine()
    eng.build_scene(name, texture_size=16, scale=float(scale) if scale else 1.0)
    ctx = capi.Context(cuda)
    ctx.upload_scene(eng.scene_desc())
    times = []
    for k in range(4):
        ctx.build_accel()
        st = ctx.stats()
        times.append(st["build_ms"])
    print("%-12s %10d triangles %9d wide nodes | build ms: first %.2f, then %s" % (spec, st["n_triangles"], st["n_bvh_nodes"], times[0], " ".join("%.2f" % t for t in times[1:])), flush=True)
    ctx.close()
    eng.close()
