import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
import numpy as np
sys.path.insert(0, %r)
from vviewer_b200 import capi
eng = capi.HostEngine()
eng.build_scene(sys.argv[1])
eng.set_render_info(width=128, height=128, samples=8, batch_size=4)
desc, rp = eng.scene_desc(), eng.render_params()
if sys.argv[2] == "1": rp.flags |= capi.PTC_FLAG_SAMPLER_SOBOL
ctx = capi.Context(capi.load_cuda()); ctx.upload_scene(desc); ctx.build_accel(); r = ctx.render(rp)
np.save(sys.argv[3], np.stack(r)); print(ctx.stats()["segments"], ctx.stats()["shadow_rays"], ctx.stats()["probe_rays"])
''' % ROOT
for scene, sob in (("Cornell", "1"), ("Cornell", "0"), ("MeshLight", "1")):
    outs = {}
    for label, env in (("plain", {"PTC_OVERLAP": "0"}), ("chunk", {"PTC_OVERLAP": "0", "PTC_MAX_SLOTS": str(128 * 128 * 2)}), ("overlap", {"PTC_OVERLAP": "5,2"})):
        e = dict(os.environ); e.update(env)
        f = "/tmp/%s_%s_%s.npy" % (scene, sob, label)
        r = subprocess.run([sys.executable, "-c", CHILD, scene, sob, f], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(scene, sob, label, r.stdout.strip()[-200:])
        import numpy as np
        outs[label] = np.load(f)
    for label in ("chunk", "overlap"):
        d = np.abs(outs[label] - outs["plain"])
        print("   %s vs plain: max %.3e  frac>1e-3 %s" % (label, d.max(), [float(np.mean(d[k].max(axis=-1) > 1e-3)) for k in range(3)]))
