import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
from oracle import loader as oracle_loader
eng = capi.HostEngine()
scene = sys.argv[1] if len(sys.argv) > 1 else "GLTF"
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if scene == "Atrium":
    eng.build_scene(scene, texture_size=64, scale=0.05)
    eng.set_render_info(width=160, height=90, samples=8, batch_size=4)
else:
    eng.build_scene(scene)
    eng.set_render_info(width=128, height=128, samples=8, batch_size=4)
desc, rp = eng.scene_desc(), eng.render_params()
rp.flags |= flags
res = {}
for label, lib in (("cuda", capi.load_cuda()), ("oracle", oracle_loader.load_oracle())):
    ctx = capi.Context(lib); ctx.upload_scene(desc); ctx.build_accel(); res[label] = ctx.render(rp); print(label, ctx.stats()["segments"]); ctx.close()
for k, nm in enumerate(["radiance", "albedo", "normal"]):
    a, b = res["cuda"][k][..., :3], res["oracle"][k][..., :3]
    d = np.abs(a - b).max(axis=-1)
    print(nm, "mean", a.mean(), b.mean(), "frac>1e-3 %.4f >1e-2 %.4f >1e-1 %.4f max %.3f" % (np.mean(d > 1e-3), np.mean(d > 1e-2), np.mean(d > 1e-1), d.max()))
    ys, xs = np.nonzero(d > 1e-2)
    for y, x in list(zip(ys, xs))[:6]:
        print("   ", y, x, a[y, x], b[y, x])
