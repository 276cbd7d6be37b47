"""Multi-GPU measurements of the product path (BASELINE configs C2-C5) on one box, through RendererPathTracing::render()
with HOST buffers: one engine drives N GPUs (NCCL communicator inside the core), the timed call includes flatten + scene upload to
every GPU + BVH build on every GPU + render + ncclReduce + readback.  Prints one JSON line per measurement and a summary table.
usage: python tools/multigpu_report.py [out.json] [--quick] [--configs=c2,c3,c4,c5] [--counts=1,8] [--verbose]
--verbose prints the host-side phases of every timed call (PTC_VERBOSE: flatten / upload / build / render / reduce / readback)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi  # noqa: E402
import torch  # noqa: E402  (device count only)

quick = "--quick" in sys.argv
verbose = "--verbose" in sys.argv
opt = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
configs = set(opt.get("configs", "c2,c3,c4,c5").split(","))
out_path = next((a for a in sys.argv[1:] if not a.startswith("--")), None)
n_gpus = torch.cuda.device_count()
rows = []


def measure(label, scene, kw, devices, split, spp=None, repeat=2):
    eng = capi.HostEngine(devices=devices if len(devices) > 1 else None)
    eng.build_scene(scene, **kw)
    if spp:
        eng.set_render_info(samples=spp)
    ri = eng.render_info()
    if len(devices) > 1:
        eng.set_render_options(split=split)
    out = [np.empty(ri["width"] * ri["height"] * 4, np.float32) for _ in range(3)]
    # warm-up call: allocations, texture / environment upload (resident afterwards), communicator.  One batch per GPU is enough for
    # that - except with the tile split, where a rank merges its small batches into work items whose size (and so the path-state
    # allocation) grows with the sample count: there the warm-up is the frame itself
    if not (len(devices) > 1 and split == "tile"):
        eng.set_render_info(samples=ri["batch_size"] * len(devices))
    eng.render_to_memory(out)
    eng.set_render_info(samples=ri["samples"])
    best = None
    for _ in range(repeat):
        if verbose:
            os.environ["PTC_VERBOSE"] = "1"
            print("[phases] %s, %d GPUs, %s" % (label, len(devices), split), file=sys.stderr, flush=True)
        t = time.perf_counter()
        eng.render_to_memory(out)
        dt = time.perf_counter() - t
        os.environ.pop("PTC_VERBOSE", None)
        st = eng.stats()
        if best is None or dt < best[0]:
            best = (dt, st)
    dt, st = best
    row = {"config": label, "scene": scene, "gpus": len(devices), "split": split if len(devices) > 1 else "none", "width": ri["width"], "height": ri["height"],
           "spp": ri["samples"], "batch": ri["batch_size"], "depth": ri["depth"], "seconds_e2e": dt, "render_ms_device": st["render_ms"], "reduce_ms": st["reduce_ms"],
           "build_ms": st["build_ms"], "segments": st["segments"], "Mseg_s_e2e": st["segments"] / dt / 1e6, "triangles": st["n_triangles"],
           "mean": float(out[0].reshape(-1, 4)[:, :3].mean()), "alpha_ok": bool(np.all(out[0].reshape(-1, 4)[:, 3] == 1.0))}
    rows.append(row)
    print(json.dumps(row), flush=True)
    eng.close()
    return row


def devs(n):
    return list(range(n))


counts = [n for n in (1, 2, 4, 8) if n <= n_gpus]
if "counts" in opt:
    counts = [n for n in counts if str(n) in opt["counts"].split(",")]
# C2: the north-star frame, strong scaling (fixed 1920x1080 x 1024 spp)
c2 = {}
for n in (counts if "c2" in configs else []):
    for split in (("sample", "tile") if n > 1 else ("none",)):
        c2[(n, split)] = measure("C2 fixed frame", "Atrium", dict(texture_size=1024), devs(n), split, spp=256 if quick else 1024)
# C3: fog at 1 and 2 GPUs (1920x1080 x 256 spp, depth 32)
for n in [c for c in counts if c <= 2 and "c3" in configs]:
    measure("C3 fog", "Fog", dict(texture_size=1024), devs(n), "sample", spp=64 if quick else None)
# C4: instanced 3840x2160 x 256 spp, tile split over 4 and 8 (and 1 for the efficiency)
for n in [c for c in counts if c in (1, 4, 8) and "c4" in configs]:
    measure("C4 instanced", "Instanced", dict(texture_size=512, scale=0.1 if quick else 1.0), devs(n), "tile", spp=32 if quick else None, repeat=1)
# C5: progressive 4096 spp, sample split over 8 (and 1)
for n in [c for c in counts if c in (1, 8) and "c5" in configs]:
    measure("C5 progressive", "Progressive", {}, devs(n), "sample", spp=256 if quick else None, repeat=1)

print("\n%-16s %4s %-7s %10s %10s %9s %9s %8s" % ("config", "gpus", "split", "s (e2e)", "Mseg/s", "efficiency", "build ms", "reduce ms"))
base = {}
for r in rows:
    key = r["config"]
    if r["gpus"] == 1:
        base[key] = r["seconds_e2e"]
    eff = base[key] / (r["seconds_e2e"] * r["gpus"]) if key in base else float("nan")
    r["efficiency_vs_1gpu"] = eff
    print("%-16s %4d %-7s %10.3f %10.1f %9.3f %9.1f %8.2f" % (key, r["gpus"], r["split"], r["seconds_e2e"], r["Mseg_s_e2e"], eff, r["build_ms"], r["reduce_ms"]))
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
