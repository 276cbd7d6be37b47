"""Small renders of scenes that cover every kernel, for compute-sanitizer (memcheck / initcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py
Round 2 adds: the two-level structure, the PMJ02BN sampler with its tables, the 48- and 63-bit Morton tiers of the hand-written
radix sort, merged batches under a tile split, a readback that crosses the staging-chunk boundary."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
cuda = capi.load_cuda()
eng = capi.HostEngine()
CASES = (("Cornell", 0, {}, {}), ("Volume5", 0, {}, {}), ("Volume3", 0, {}, {}), ("Transparency", capi.PTC_FLAG_SAMPLER_SOBOL, {}, {}),
         ("EnvironmentMapPBR00", capi.PTC_FLAG_ENV_IMPORTANCE, {}, {}), ("MeshLight", 0, {}, {}),
         ("Instanced", 0, {"scale": 0.004, "texture_size": 32}, {}), ("Progressive", 0, {"scale": 0.05}, {}), ("Fog", 0, {"scale": 0.02, "texture_size": 16}, {}),
         # round 2
         ("Instanced", 0, {"scale": 0.004, "texture_size": 32}, {"accel": capi.PTC_ACCEL_TWO_LEVEL}),
         ("Cornell", capi.PTC_FLAG_SAMPLER_PMJ, {}, {"pmj": True}),
         ("Atrium", 0, {"scale": 0.02, "texture_size": 32}, {"morton": "16"}),
         ("Atrium", 0, {"scale": 0.02, "texture_size": 32}, {"morton": "21", "merge": True}),
         ("MeshLight", 0, {}, {"big": True}))
for scene, flags, kw, opt in CASES:
    eng.build_scene(scene, **kw)
    if opt.get("big"):
        eng.set_render_info(width=1056, height=1000, samples=1, batch_size=1, depth=2)  # 16.9 MB per image: two staging chunks
    else:
        eng.set_render_info(width=48, height=40, samples=32 if opt.get("merge") else 8, batch_size=4, depth=6)
    rp = eng.render_params()
    rp.flags |= flags
    if "morton" in opt:
        os.environ["PTC_MORTON_BITS"] = opt["morton"]
    ctx = capi.Context(cuda)
    if "accel" in opt:
        ctx.set_accel_mode(opt["accel"])
    if opt.get("pmj"):
        ctx.set_sampler_tables()
    ctx.upload_scene(eng.scene_desc())
    ctx.build_accel()
    os.environ.pop("PTC_MORTON_BITS", None)
    ctx.render(rp)
    rp.split_mode, rp.rank, rp.world = capi.PTC_SPLIT_TILE, 1, 3
    rp.tile_size = 8
    ctx.render(rp)
    print(scene, sorted(opt), ctx.stats()["segments"], flush=True)
    ctx.close()
print("done")
