"""Short single-GPU run of the bench workload for ncu (never a bench number): N batches of the C2 atrium."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
batches = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scene = sys.argv[2] if len(sys.argv) > 2 else "Atrium"
eng = capi.HostEngine()
eng.build_scene(scene, texture_size=1024)
ctx = capi.Context(capi.load_cuda())
ctx.upload_scene(eng.scene_desc())
ctx.build_accel()
rp = eng.render_params()
rp.samples = batches * rp.batch_size
ctx.render(rp, want_aovs=False)
print(ctx.stats())
