import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from vviewer_b200 import capi
from test_gpu_parity import ray_set, both
eng = capi.HostEngine()
# --- cornell rays
eng.build_scene("Cornell")
cu, orc = both(capi, eng.scene_desc())
rng = np.random.default_rng(11)
box = 1.0
rays = np.concatenate([ray_set(rng, 30000, -box, box), ray_set(rng, 30000, -box, box, aim=(-box / 4, box / 4))])
ia, pa, ta, ua, va = cu.trace_closest(rays)
ib, pb, tb, ub, vb = orc.trace_closest(rays)
same = (ia == ib) & (pa == pb)
bad = np.where(~same)[0]
print("cornell mismatches", len(bad))
for k in bad[:10]:
    print(k, rays[k], "cuda", ia[k], pa[k], ta[k], ua[k], va[k], "orc", ib[k], pb[k], tb[k], ub[k], vb[k])
hit = same & (ib >= 0)
db = np.maximum(np.abs(ua[hit] - ub[hit]), np.abs(va[hit] - vb[hit]))
print("bary p99.9", np.percentile(db, 99.9), db.max(), "dt max", np.abs(ta[hit]-tb[hit]).max())
# --- bsdf
cu2, orc2 = capi.Context(capi.load_cuda()), capi.Context(capi.load_oracle())
rng = np.random.default_rng(5)
n = 100000
def dirs():
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:, 1] = np.abs(d[:, 1])
    return d
wi, wo = dirs(), dirs()
wi[: n // 20, 1] *= -1
params = np.stack([rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(0.035, 1, n)], axis=1).astype(np.float32)
fa, pa_ = cu2.bsdf_eval(params, wi, wo)
fb, pb_ = orc2.bsdf_eval(params, wi, wo)
rel = np.abs(fa - fb) / np.maximum(np.abs(fb), 1e-3)
relp = np.abs(pa_ - pb_) / np.maximum(np.abs(pb_), 1e-3)
print("eval rel max", rel.max(), "frac>1e-5", (rel.max(axis=1) > 1e-5).mean(), "pdf rel max", relp.max(), "frac>1e-5", (relp > 1e-5).mean())
w = np.argsort(-rel.max(axis=1))[:5]
for k in w:
    print(params[k], wi[k], wo[k], fa[k], fb[k])
u = rng.uniform(0, 1, (n, 3)).astype(np.float32)
wa, fa, pa_ = cu2.bsdf_sample(params, wo, u)
wb, fb, pb_ = orc2.bsdf_sample(params, wo, u)
ok = pb_ >= 1e-6
print("sample wi max diff", np.abs(wa - wb)[ok].max(), "pdf rel", (np.abs(pa_ - pb_)[ok] / np.maximum(pb_[ok], 1e-3)).max(), "zero mismatch", np.mean((pa_ < 1e-6) != (pb_ < 1e-6)))
