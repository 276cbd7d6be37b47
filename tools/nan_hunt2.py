import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
eng = capi.HostEngine(); eng.build_scene("MeshLight"); ri = eng.render_info(); B = ri["batch_size"]
desc = eng.scene_desc()
ctx = capi.Context(capi.load_cuda()); ctx.upload_scene(desc); ctx.build_accel()
def bad(nb):
    eng.set_render_info(samples=nb * B); rp = eng.render_params()
    r = ctx.render(rp)[0]
    return not np.isfinite(r[242, 92]).all()
lo, hi = 0, 2048 // B
assert bad(hi)
while hi - lo > 1:
    mid = (lo + hi) // 2
    if bad(mid): hi = mid
    else: lo = mid
print("batch size", B, "first bad batch count", hi, "-> sample indices", (hi - 1) * B, "..", hi * B - 1)
# depth scan: which max depth first shows it
for depth in range(1, 10):
    eng.set_render_info(samples=hi * B, depth=depth); rp = eng.render_params()
    r = ctx.render(rp)[0]
    print("depth", depth, "finite", bool(np.isfinite(r[242, 92]).all()), r[242, 92])
