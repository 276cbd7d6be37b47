"""Small renders of scenes that cover every kernel, for compute-sanitizer (memcheck / initcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
cuda = capi.load_cuda()
eng = capi.HostEngine()
for scene, flags, kw in (("Cornell", 0, {}), ("Volume5", 0, {}), ("Volume3", 0, {}), ("Transparency", capi.PTC_FLAG_SAMPLER_SOBOL, {}),
                         ("EnvironmentMapPBR00", capi.PTC_FLAG_ENV_IMPORTANCE, {}), ("MeshLight", 0, {}),
                         ("Instanced", 0, {"scale": 0.004, "texture_size": 32}), ("Progressive", 0, {"scale": 0.05}), ("Fog", 0, {"scale": 0.02, "texture_size": 16})):
    eng.build_scene(scene, **kw)
    eng.set_render_info(width=48, height=40, samples=8, batch_size=4, depth=6)
    rp = eng.render_params()
    rp.flags |= flags
    ctx = capi.Context(cuda)
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    ctx.render(rp)
    rp.split_mode, rp.rank, rp.world = capi.PTC_SPLIT_TILE, 1, 3
    ctx.render(rp)
    print(scene, ctx.stats()["segments"], flush=True)
    ctx.close()
print("done")
