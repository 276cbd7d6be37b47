"""Counts non-finite pixels of a recipe rendered at a high sample count (diagnosis).  usage: python tools/nan_hunt.py scene spp"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi
scene, spp = sys.argv[1], int(sys.argv[2])
eng = capi.HostEngine(); eng.build_scene(scene); ri = eng.render_info()
eng.set_render_info(samples=spp)
desc, rp = eng.scene_desc(), eng.render_params()
for label, env in (("default", {}), ("no fusion", {"PTC_NO_PROBE_FUSION": "1"})):
    os.environ.update(env)
    ctx = capi.Context(capi.load_cuda()); ctx.upload_scene(desc); ctx.build_accel()
    r = ctx.render(rp)[0]
    bad = ~np.isfinite(r).all(axis=-1)
    print(scene, spp, label, "non-finite pixels:", int(bad.sum()), list(zip(*np.nonzero(bad)))[:5], "mean", float(np.nanmean(r[..., :3])))
    # which batch: render batches separately
    if bad.any():
        rp2 = eng.render_params(); rp2.samples = rp.batch_size
        # not possible to select a batch index through the API: report only
    ctx.close()
    for k in env: os.environ.pop(k, None)
