"""Runs the BASELINE.json workload configs for a few batches on one GPU and prints build time,
throughput and the per-kernel time split.  Not a bench line (PTC_FLAG_TIME_KERNELS serialises launches).
usage: python tools/workloads.py [name[:scale[:batches]] ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
specs = sys.argv[1:] or ["Cornell", "Atrium", "Fog", "Progressive", "Instanced:0.05"]
cuda = capi.load_cuda()
for spec in specs:
    parts = spec.split(":")
    name = parts[0]
    scale = float(parts[1]) if len(parts) > 1 else 1.0
    batches = int(parts[2]) if len(parts) > 2 else 2
    eng = capi.HostEngine()
    t = time.time()
    eng.build_scene(name, texture_size=1024 if name in ("Atrium", "Fog") else 512, scale=scale)
    t_host = time.time() - t
    ctx = capi.Context(cuda)
    t = time.time(); ctx.upload_scene(eng.scene_desc()); t_up = time.time() - t
    ctx.build_accel()
    b = ctx.stats()
    rp = eng.render_params()
    ri = eng.render_info()
    for flags in (0, capi.PTC_FLAG_TIME_KERNELS):
        rp.samples = batches * rp.batch_size
        rp.flags = flags
        ctx.render(rp, want_aovs=False)
        st = ctx.stats()
        if flags == 0:
            plain = st
    spp_frame = ri["samples"]
    print("%-12s scale %.2f  %dx%d batch %d depth %d | tris %d wide nodes %d | host %.1fs upload %.2fs build %.1f ms | %.1f Mseg/s  %.1f ms/batch  %.2f s/frame(%d spp)"
          % (name, scale, ri["width"], ri["height"], ri["batch_size"], ri["depth"], b["n_triangles"], b["n_bvh_nodes"], t_host, t_up, b["build_ms"],
             plain["segments"] / plain["render_ms"] / 1e3, plain["render_ms"] / batches, plain["render_ms"] / batches * (spp_frame // ri["batch_size"]) / 1e3, spp_frame))
    print("             seg/path %.2f  shadow rays %d (hops %d)  probe rays %d (hops %d) | timed: extend %.1f  shade %.1f  shadow+probe %.1f ms of %.1f"
          % (st["segments"] / (ri["width"] * ri["height"] * rp.samples), st["shadow_rays"], st["shadow_hops"], st["probe_rays"], st["probe_hops"],
             st["trace_ms"], st["shade_ms"], st["shadow_ms"], st["render_ms"]))
    ctx.close(); eng.close()
