"""Where the end-to-end time of one RendererPathTracing::render() call goes (PTC_VERBOSE=1 prints the phases).
usage: PTC_VERBOSE=1 python tools/e2e_breakdown.py [scene] [batches]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "Atrium"
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = capi.HostEngine()
eng.build_scene(scene)
ri = eng.render_info()
eng.set_render_info(samples=batches * ri["batch_size"])
for i in range(3):
    t = time.perf_counter()
    eng.render_to_memory()
    dt = time.perf_counter() - t
    st = eng.stats()
    print("call %d: %.1f ms wall, render_ms %.1f, %.1f Mseg/s end to end, %.1f device" % (i, dt * 1e3, st["render_ms"], st["segments"] / dt / 1e6,
                                                                                       st["segments"] / st["render_ms"] / 1e3), flush=True)
