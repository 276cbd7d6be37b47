"""Diagnosis (not a bench): per-ray node visits / triangle tests and warp utilisation of k_extend, from the
-DPTC_TRAV_STATS build (`make stats`).  usage: python tools/trav_stats.py [batches] [scene]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
batches = int(sys.argv[1]) if len(sys.argv) > 1 else 1
scene = sys.argv[2] if len(sys.argv) > 2 else "Atrium"
eng = capi.HostEngine()
eng.build_scene(scene, texture_size=256)
lib = capi.load_ptc(os.environ.get("PTC_LIB") or os.path.join(capi.LIB_DIR, "libptc_cuda_stats.so"))
ctx = capi.Context(lib)
ctx.upload_scene(eng.scene_desc())
ctx.build_accel()
rp = eng.render_params()
rp.samples = batches * rp.batch_size
rp.flags = capi.PTC_FLAG_TIME_KERNELS
ctx.render(rp, want_aovs=False)
st = ctx.stats()
rays = st["segments"]
nv, tt, ni, ti = st["reserved"]
print("rays %d  nodes/ray %.2f  tris/ray %.2f  node iters %d (%.1f lanes)  tri iters %d (%.1f lanes)  trace_ms %.2f  shade_ms %.2f  wide nodes %d" % (
    rays, nv / rays, tt / rays, ni, nv / max(ni, 1), ti, tt / max(ti, 1), st["trace_ms"], st["shade_ms"], st["n_bvh_nodes"]))
