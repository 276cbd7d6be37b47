"""Pins the CPU oracle against the reference's 31 golden renders (assets/unittests/*_ref.hdr, copied to
tests/golden/reference_images/): renders every RenderTests.cpp recipe with the oracle at the golden's own sample count
and writes tests/golden/oracle_vs_reference.json (committed).  Run here on CPU:

    python tests/golden/run_oracle_goldens.py [--spp-scale 1.0] [names...]
    python tests/golden/run_oracle_goldens.py --backend cuda [--spp-scale 4.0] [names...]   (on a B200: writes cuda_vs_reference.json)

Metrics: RGB MSE after the same RGBE quantisation the goldens went through, mean-luminance ratio,
99th percentile relative error after a 3x3 box filter.  GLTF_ref goes through the from-scratch glTF importer
(vviewer_b200/host/io_gltf.cpp); Denoise_ref needs OIDN (out of scope): Denoise_ref_{radiance,albedo,normal} are checked instead."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from imgmetrics import mean_lum_ratio, mse, p99_rel_err, rgbe_roundtrip  # noqa: E402
from vviewer_b200 import capi  # noqa: E402

SCENES = ["FurnacePBR", "FurnaceLambert", "EnvironmentMap", "EnvironmentMapPBR00", "EnvironmentMapPBR01", "EnvironmentMapPBR10",
          "EnvironmentMapPBR11", "EnvironmentMapLambert", "Volume0", "Volume1", "Volume2", "Volume3", "Volume4", "Volume5", "Volume6",
          "Volume7", "Volume8", "Volume9", "PointLight", "DirectionalLight", "MeshLight", "Transparency", "NormalMap", "GLTF", "Hierarchy",
          "DepthOfField", "SharedComponents", "Denoise"]


def compare(img, ref_path):
    ref = capi.read_hdr(ref_path)
    q = rgbe_roundtrip(img)
    return {"mse": mse(q, ref), "lum_ratio": mean_lum_ratio(q, ref), "p99_rel": p99_rel_err(q, ref)}


def main():
    args = sys.argv[1:]
    scale = 1.0
    backend = "oracle"
    while args and args[0] in ("--spp-scale", "--backend"):
        if args[0] == "--spp-scale":
            scale = float(args[1])
        else:
            backend = args[1]
        args = args[2:]
    names = args or SCENES
    out_path = os.path.join(HERE, "oracle_vs_reference.json" if backend == "oracle" else "cuda_vs_reference.json")
    results = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for name in names:
        eng = capi.HostEngine() if backend == "oracle" else capi.HostEngine()
        eng.build_scene(name)
        ri = eng.render_info()
        spp = max(ri["batch_size"], int(ri["samples"] * scale) // ri["batch_size"] * ri["batch_size"])
        eng.set_render_info(samples=spp)
        t = time.time()
        rad, alb, nrm = eng.render_to_memory()
        dt = time.time() - t
        ref_dir = os.path.join(HERE, "reference_images")
        if name == "Denoise":
            r = {"radiance": compare(rad, os.path.join(ref_dir, "Denoise_ref_radiance.hdr")),
                 "albedo": compare(alb, os.path.join(ref_dir, "Denoise_ref_albedo.hdr")),
                 "normal": compare(nrm, os.path.join(ref_dir, "Denoise_ref_normal.hdr"))}
            r = {"mse": r["radiance"]["mse"], "lum_ratio": r["radiance"]["lum_ratio"], "p99_rel": r["radiance"]["p99_rel"], "aov": r}
        else:
            r = compare(rad, os.path.join(ref_dir, name + "_ref.hdr"))
        r.update({"spp": spp, "seconds": round(dt, 1), "segments": eng.stats()["segments"]})
        results[name] = r
        print("%-24s spp %5d %6.1fs mse %.3e lum %.4f p99 %.4f" % (name, spp, dt, r["mse"], r["lum_ratio"], r["p99_rel"]), flush=True)
        eng.close()
        json.dump(results, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
