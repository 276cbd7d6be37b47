import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
eng = capi.HostEngine(); eng.build_scene("MeshLight"); ri = eng.render_info(); B = ri["batch_size"]
lib = capi.load_ptc(os.environ["PTC_LIB"])
ctx = capi.Context(lib); ctx.upload_scene(eng.scene_desc()); ctx.build_accel()
eng.set_render_info(samples=26 * B, depth=2); rp = eng.render_params()
r = ctx.render(rp)[0]
print("pixel", r[242, 92], "bad pixels", int((~np.isfinite(r).all(axis=-1)).sum()))
