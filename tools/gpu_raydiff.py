"""Bring-up helper (not a test): ray-set mismatches between the CUDA traversal and the oracle, with details."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from vviewer_b200 import capi
from test_gpu_parity import ray_set, both
eng = capi.HostEngine()
cases = [("EnvironmentMap", {}, 4.0), ("Hierarchy", {}, 12.0), ("SharedComponents", {}, 60.0), ("Cornell", {}, 1.0), ("Atrium", dict(texture_size=4, scale=0.25), 8.0)]
for scene, kw, box in cases:
    eng.build_scene(scene, **kw)
    cu, orc = both(capi, eng.scene_desc())
    rng = np.random.default_rng(11)
    rays = np.concatenate([ray_set(rng, 30000, -box, box), ray_set(rng, 30000, -box, box, aim=(-box / 4, box / 4))])
    ia, pa, ta, ua, va = cu.trace_closest(rays)
    ib, pb, tb, ub, vb = orc.trace_closest(rays)
    same = (ia == ib) & (pa == pb)
    bad = np.where(~same)[0]
    dt = np.abs(ta - tb)
    missed = bad[(ta[bad] > tb[bad] * (1 + 1e-5))]
    extra = bad[(ta[bad] < tb[bad] * (1 - 1e-5))]
    print("%-18s mismatches %d  ties %d  cuda-missed %d  cuda-nearer %d" % (scene, len(bad), len(bad) - len(missed) - len(extra), len(missed), len(extra)))
    for k in list(missed[:6]) + list(extra[:4]):
        print("   ray", k, rays[k].tolist(), "cuda", ia[k], pa[k], ta[k], "orc", ib[k], pb[k], tb[k], ub[k], vb[k])
    cu.close(); orc.close()
