"""Cost and benefit of PTC_FLAG_ENV_IMPORTANCE on the bench scene: throughput with / without, and the error of a 16-spp render of each
against a 1024-spp reference (quarter resolution).  usage: python tools/env_importance_run.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
eng = capi.HostEngine(); eng.build_scene("Atrium")
ctx = capi.Context(capi.load_cuda()); ctx.upload_scene(eng.scene_desc()); ctx.build_accel()
rp = eng.render_params()
for label, fl in (("plain", 0), ("env-importance", capi.PTC_FLAG_ENV_IMPORTANCE)):
    rp.flags = fl; rp.samples = 4 * rp.batch_size
    ctx.render(rp, want_aovs=False); ctx.render(rp, want_aovs=False); st = ctx.stats()
    print("%-15s %.1f Mseg/s, %.2f ms/batch, shadow rays per segment %.3f" % (label, st["segments"] / st["render_ms"] / 1e3, st["render_ms"] / 4, st["shadow_rays"] / st["segments"]))
rp.width //= 4; rp.height //= 4
imgs = {}
for label, fl, spp in (("ref", capi.PTC_FLAG_ENV_IMPORTANCE, 4096), ("ref_plain", 0, 4096), ("plain16", 0, 16), ("imp16", capi.PTC_FLAG_ENV_IMPORTANCE, 16)):
    rp.flags = fl; rp.samples = spp; rp.batch_size = 16
    imgs[label] = ctx.render(rp, want_aovs=False)[0][..., :3]
ref = imgs["ref"]
print("mean: flag %.5f plain %.5f ratio %.4f" % (ref.mean(), imgs["ref_plain"].mean(), ref.mean() / imgs["ref_plain"].mean()))
for k in ("plain16", "imp16"):
    print("%-8s 16 spp: MSE vs its own 4096-spp reference %.4e" % (k, np.mean((imgs[k] - (imgs["ref_plain"] if k == "plain16" else ref)) ** 2)))
