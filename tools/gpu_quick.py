"""Quick GPU bring-up check (not a test): CUDA vs oracle on a few recipe scenes."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
from oracle import loader as oracle_loader

names = sys.argv[1:] or ["FurnaceLambert", "EnvironmentMap", "MeshLight", "Volume5", "Transparency"]
cuda = capi.load_cuda()
orc = oracle_loader.load_oracle()
eng = capi.HostEngine()
print("backend ok:", eng.backend_ok(), eng.last_error())
for name in names:
    eng.build_scene(name)
    eng.set_render_info(samples=16, batch_size=16)
    desc = eng.scene_desc()
    rp = eng.render_params()
    res = {}
    for label, lib in (("cuda", cuda), ("oracle", orc)):
        ctx = capi.Context(lib)
        t = time.time(); ctx.upload_scene(desc); t_up = time.time() - t
        t = time.time(); ctx.build_accel(); t_b = time.time() - t
        t = time.time(); rad, alb, nrm = ctx.render(rp); t_r = time.time() - t
        st = ctx.stats()
        res[label] = (rad, alb, nrm, st, ctx)
        print("%-16s %-6s upload %.3fs build %.3fs render %.3fs seg %d sh %d/%d pr %d/%d render_ms %.1f" % (
            name, label, t_up, t_b, t_r, st["segments"], st["shadow_rays"], st["shadow_hops"], st["probe_rays"], st["probe_hops"], st["render_ms"]))
    a, b = res["cuda"], res["oracle"]
    for k, nm in enumerate(["radiance", "albedo", "normal"]):
        d = np.abs(a[k][..., :3] - b[k][..., :3])
        print("   %-8s mean cuda %.6f oracle %.6f  mse %.3e  maxabs %.3e  frac>1e-3 %.5f" % (
            nm, a[k][..., :3].mean(), b[k][..., :3].mean(), np.mean(d ** 2), d.max(), np.mean(d.max(axis=-1) > 1e-3)))
    # LBVH parity
    la, lb = a[4].get_lbvh(), b[4].get_lbvh()
    ok = all(np.array_equal(la[k], lb[k]) for k in ("morton", "order", "parent", "left", "right", "aabb"))
    print("   lbvh n=%d bit-exact=%s" % (la["n"], ok))
    if not ok:
        for k in ("morton", "order", "parent", "left", "right", "aabb"):
            print("     ", k, np.array_equal(la[k], lb[k]), int(np.sum(la[k] != lb[k])))
    # ray parity
    rng = np.random.default_rng(1)
    n = 20000
    o = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.full((n, 1), 1e-3, np.float32), d, np.full((n, 1), 1e4, np.float32)], axis=1)
    ia, pa, ta, ua, va = a[4].trace_closest(rays)
    ib, pb, tb, ub, vb = b[4].trace_closest(rays)
    same = (ia == ib) & (pa == pb)
    print("   rays: hit frac %.3f id mismatch %d  max|dt| %.3e" % (np.mean(ib >= 0), int(np.sum(~same)), float(np.max(np.abs(ta - tb)[same])) if same.any() else -1))
    for x in res.values():
        x[4].close()
