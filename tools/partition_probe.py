"""How well does a hand partition of one render balance?  Renders every rank's share of a tile / sample split on ONE GPU and prints
the device time of each share next to the full render's (diagnosis for the multi-GPU efficiency of a config).
usage: python tools/partition_probe.py Scene[:scale] world [spp]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from vviewer_b200 import capi  # noqa: E402
import partition_util  # noqa: E402

name, _, scale = sys.argv[1].partition(":")
world = int(sys.argv[2])
eng = capi.HostEngine()
eng.build_scene(name, texture_size=512, scale=float(scale) if scale else 1.0)
if len(sys.argv) > 3:
    eng.set_render_info(samples=int(sys.argv[3]))
ctx = capi.Context(capi.load_cuda())
ctx.upload_scene(eng.scene_desc())
ctx.build_accel()
rp = eng.render_params()
ctx.render(rp, want_aovs=False)
ctx.render(rp, want_aovs=False)
full = ctx.stats()
print("%s full: %.1f ms, %d segments, %.1f Mseg/s" % (sys.argv[1], full["render_ms"], full["segments"], full["segments"] / full["render_ms"] / 1e3))
for mode in ("tile", "sample"):
    ms, seg = [], []
    for r in range(world):
        p = partition_util.partition(eng.render_params(), r, world, mode, tile_size=32)
        ctx.render(p, want_aovs=False)
        st = ctx.stats()
        ms.append(st["render_ms"])
        seg.append(st["segments"])
    print("  %-6s shares ms: %s | max %.1f vs full/world %.1f (balance %.3f) | segments max/mean %.3f | Mseg/s of the slowest share %.1f" % (
        mode, " ".join("%.0f" % m for m in ms), max(ms), full["render_ms"] / world, full["render_ms"] / world / max(ms), max(seg) / (sum(seg) / world),
        seg[ms.index(max(ms))] / max(ms) / 1e3))
