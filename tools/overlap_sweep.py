"""Sweep of the two-wavefront overlap configuration (PTC_OVERLAP = traceBlocksPerSM,shadeBlocksPerSM; 0 = off).
usage: python tools/overlap_sweep.py [scene] [batches] cfg cfg ...     (each cfg runs in a fresh process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r)
from vviewer_b200 import capi
scene, batches = sys.argv[1], int(sys.argv[2])
eng = capi.HostEngine()
eng.build_scene(scene)
ctx = capi.Context(capi.load_ptc(os.environ["PTC_LIB"]) if os.environ.get("PTC_LIB") else capi.load_cuda())  # PTC_LIB: experimental build variants (tools only)
ctx.upload_scene(eng.scene_desc()); ctx.build_accel()
rp = eng.render_params()
rp.samples = 3 * rp.batch_size
ctx.render(rp, want_aovs=False)
rp.samples = batches * rp.batch_size
best = 0
for rep in range(2):
    ctx.render(rp, want_aovs=False)
    st = ctx.stats()
    best = max(best, st["segments"] / st["render_ms"] / 1e3)
print(os.path.basename(os.environ.get("PTC_LIB", "")), "%%-8s %%-10s %%8.1f Mseg/s  %%7.2f ms/batch" %% (os.environ.get("PTC_OVERLAP", "default"), scene, best, st["segments"] / best / 1e3 / batches))
''' % ROOT

scene = sys.argv[1] if len(sys.argv) > 1 else "Atrium"
batches = sys.argv[2] if len(sys.argv) > 2 else "8"
for cfg in sys.argv[3:] or ["0", "5,2", "6,1", "4,3", "4,2", "6,2"]:
    env = dict(os.environ)
    if cfg == "default":
        env.pop("PTC_OVERLAP", None)
    else:
        env["PTC_OVERLAP"] = cfg
    r = subprocess.run([sys.executable, "-c", CHILD, scene, batches], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "(no output)", flush=True)
