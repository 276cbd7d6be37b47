"""Evidence for the four firefly goldens (Volume4/5/8/9: a point light or an emissive plane INSIDE a scattering medium).
Renders each recipe on the GPU at 8x the golden's samples (">= 8x") and splits the squared error against the reference
image into (a) the golden's own isolated spikes (imgmetrics.firefly_mask: luminance > 3x the 3x3 median) and (b) everything else.
Also renders the same recipe twice at the golden's OWN sample count with two different seeds... (not possible: the sampler is
keyed by pixel and sample index) - instead the product's own 1x render is compared with its 8x render: the same spikes must
appear in a 1x render of ours if they are Monte-Carlo noise of this estimator.
usage: python tools/golden_fireflies.py [out.json]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from vviewer_b200 import capi  # noqa: E402
from imgmetrics import firefly_mask, mean_lum_ratio, mse, mse_masked, p99_rel_err, rgbe_roundtrip  # noqa: E402

REF = os.path.join(ROOT, "tests", "golden", "reference_images")
out = {}
eng = capi.HostEngine()
for scene in sys.argv[2:] or ["Volume4", "Volume5", "Volume8", "Volume9", "Volume3", "MeshLight"]:
    eng.build_scene(scene)
    spp = eng.render_info()["samples"]
    ref = capi.read_hdr(os.path.join(REF, scene + "_ref.hdr"))
    row = {"golden_spp": spp}
    imgs = {}
    for mult in (1, 8):
        eng.set_render_info(samples=mult * spp)
        imgs[mult] = rgbe_roundtrip(eng.render_to_memory()[0])
    m = firefly_mask(ref)
    row["golden_firefly_pixels"] = int(m.sum())
    row["own_1x_firefly_pixels"] = int(firefly_mask(imgs[1]).sum())
    row["own_8x_firefly_pixels"] = int(firefly_mask(imgs[8]).sum())
    for mult in (1, 8):
        row["mse_%dx" % mult] = mse(imgs[mult], ref)
        row["mse_%dx_without_golden_fireflies" % mult] = mse_masked(imgs[mult], ref, m)
        row["lum_%dx" % mult] = mean_lum_ratio(imgs[mult], ref)
        row["p99_%dx" % mult] = p99_rel_err(imgs[mult], ref)
    # share of the squared error that sits in the golden's spikes
    d = ((imgs[8][..., :3].astype(np.float64) - ref[..., :3].astype(np.float64)) ** 2).sum(axis=-1)
    row["error_share_in_golden_fireflies_8x"] = float(d[m].sum() / max(d.sum(), 1e-30))
    # our own 1x render against our own 8x render: the noise floor of THIS estimator at the golden's sample count
    row["own_1x_vs_own_8x_mse"] = mse(imgs[1], imgs[8])
    out[scene] = row
    print(scene, json.dumps(row), flush=True)
eng.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1, sort_keys=True)
