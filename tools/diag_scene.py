"""CUDA-vs-oracle difference breakdown of one recipe (diagnosis, not a test).
usage: python tools/diag_scene.py Scene [scale] [texsize] [w h spp batch]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vviewer_b200 import capi  # noqa: E402
from oracle import loader as oracle_loader  # noqa: E402

name = sys.argv[1]
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
tex = int(sys.argv[3]) if len(sys.argv) > 3 else 64
w, h, spp, batch = (int(x) for x in sys.argv[4:8]) if len(sys.argv) > 7 else (128, 72, 8, 4)
eng = capi.HostEngine()
eng.build_scene(name, texture_size=tex, scale=scale)
eng.set_render_info(width=w, height=h, samples=spp, batch_size=batch)
desc = eng.scene_desc()
for depth in (1, 2, 4, eng.render_info()["depth"]):
    rp = eng.render_params()
    rp.depth = depth
    res = {}
    for label, lib in (("cuda", capi.load_cuda()), ("oracle", oracle_loader.load_oracle())):
        ctx = capi.Context(lib)
        ctx.upload_scene(desc)
        ctx.build_accel()
        res[label] = ctx.render(rp) + (ctx.stats(),)
        ctx.close()
    a, b = res["cuda"], res["oracle"]
    print("== %s scale %g tex %d depth %d: segments cuda %d oracle %d" % (name, scale, tex, depth, a[3]["segments"], b[3]["segments"]))
    for k, nm in enumerate(["radiance", "albedo", "normal"]):
        d = np.abs(a[k][..., :3] - b[k][..., :3]).max(axis=-1)
        rel = d / np.maximum(b[k][..., :3].max(axis=-1), 1e-2)
        print("   %-8s mean cuda %.6f oracle %.6f | frac>1e-3 %.4f  >1e-2 %.4f  >1e-1 %.4f | max %.3e  median %.3e | rel: median %.2e p90 %.2e p99 %.2e frac>2%% %.4f" % (
            nm, a[k][..., :3].mean(), b[k][..., :3].mean(), np.mean(d > 1e-3), np.mean(d > 1e-2), np.mean(d > 1e-1), d.max(), np.median(d),
            np.median(rel), np.percentile(rel, 90), np.percentile(rel, 99), np.mean(rel > 0.02)))
