"""Runs one workload under several environment settings, each in a fresh process (the library reads its tuning variables at
ptc_create), and prints throughput + the per-kernel split.  Not a bench line.
usage: python tools/env_sweep.py Scene[:scale[:batches]] "VAR=a,VAR2=b" "VAR=c" ...   ("-" = no variables)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
from vviewer_b200 import capi
name, scale, batches = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
lib = capi.load_ptc(os.environ["PTC_LIB"]) if os.environ.get("PTC_LIB") else capi.load_cuda()
eng = capi.HostEngine()
eng.build_scene(name, texture_size=1024 if name in ("Atrium", "Fog") else 512, scale=scale)
ctx = capi.Context(lib)
ctx.upload_scene(eng.scene_desc()); ctx.build_accel()
b = ctx.stats()
rp = eng.render_params()
out = {"build_ms": b["build_ms"], "tris": b["n_triangles"]}
for flags in (0, 0, capi.PTC_FLAG_TIME_KERNELS):
    rp.samples = batches * rp.batch_size; rp.flags = flags
    img = ctx.render(rp, want_aovs=False)
    st = ctx.stats()
    if flags == 0:
        out["Mseg_s"] = st["segments"] / st["render_ms"] / 1e3; out["ms_batch"] = st["render_ms"] / batches; out["mean"] = float(img[..., :3].mean())
    else:
        out.update({"extend": st["trace_ms"] / batches, "shade": st["shade_ms"] / batches, "chains": st["shadow_ms"] / batches, "bin": 0.0,
                    "timed_total": st["render_ms"] / batches, "reserved": st["reserved"], "segments": st["segments"]})
print(json.dumps(out))
''' % ROOT

spec = sys.argv[1].split(":")
name, scale, batches = spec[0], (spec[1] if len(spec) > 1 else "1.0"), (spec[2] if len(spec) > 2 else "4")
for setting in sys.argv[2:] or ["-"]:
    env = dict(os.environ)
    if setting != "-":
        for kv in setting.split(","):
            k, v = kv.split("=", 1)
            env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD, name, scale, batches], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        print("%-40s FAILED: %s" % (setting, r.stderr[-400:]))
        continue
    o = json.loads(r.stdout.strip().splitlines()[-1])
    extra = ""
    if any(o["reserved"]):
        nv, tt, ni, ti = o["reserved"]
        extra = " | nodes/ray %.2f tris/ray %.2f lanes/node-iter %.1f lanes/tri-iter %.1f" % (nv / o["segments"], tt / o["segments"], nv / max(ni, 1), tt / max(ti, 1))
    print("%-40s %8.1f Mseg/s %7.2f ms/batch | timed: extend %.2f shade %.2f chains %.2f bin %.2f of %.2f | mean %.6f build %.1f ms%s" % (
        setting, o["Mseg_s"], o["ms_batch"], o["extend"], o["shade"], o["chains"], o["bin"], o["timed_total"], o["mean"], o["build_ms"], extra), flush=True)
