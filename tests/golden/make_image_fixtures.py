"""Generates tests/golden/stb_decode_fixtures.json with the REFERENCE's own image decoder.

The reference decodes every 8-bit texture with stbi_load / stbi_load_from_memory(..., STBI_rgb_alpha) and every HDR map
with stbi_loadf, after a global stbi_set_flip_vertically_on_load(true) (core/Image.cpp:9-43, io/AssimpLoadModel.cpp:190,
211-223).  oracle/_ref/stb_decode (`make oracle-ref`) is that decoder compiled from the vendored sources under
/root/reference.  It cannot travel, so this script records, for every image the bundled assets contain (the three PNG
textures, the five JPEGs embedded in DamagedHelmet.gltf, harbor.hdr), the size, channel count, SHA-256 of the decoded
RGBA8 (or RGBA32F) rows and a few probe texels.  tests/test_import.py holds the from-scratch decoders of
vviewer_b200/host/io_image.cpp / io_jpeg.cpp to these numbers bit for bit.

    make oracle-ref && python tests/golden/make_image_fixtures.py
"""
import base64
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
TOOL = os.path.join(ROOT, "oracle", "_ref", "stb_decode")


def probes(h, w):
    rng = np.random.default_rng(12345)
    return [(int(rng.integers(0, h)), int(rng.integers(0, w))) for _ in range(16)]


def run(path, hdr, flip):
    cmd = [TOOL] + (["--hdr"] if hdr else []) + [path] + (["flip"] if flip else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, check=True).stdout
    nl = out.index(b"\n")
    w, h, c = [int(x) for x in out[:nl].split()]
    raw = out[nl + 1:]
    arr = np.frombuffer(raw, np.float32 if hdr else np.uint8).reshape(h, w, 4)
    return w, h, c, raw, arr


def embedded_images(gltf_path):
    g = json.load(open(gltf_path))
    for i, im in enumerate(g["images"]):
        uri = im["uri"]
        assert uri.startswith("data:")
        yield i, base64.b64decode(uri[uri.index(",") + 1:])


def main():
    if not os.path.exists(TOOL):
        sys.exit("build oracle/_ref/stb_decode first: make oracle-ref")
    fx = {}
    jobs = [("assets/textures/%s.png" % n, False) for n in ("checkerboard", "normal", "circular_gradient")]
    jobs.append(("assets/HDR/harbor.hdr", True))
    with tempfile.TemporaryDirectory() as tmp:
        for i, data in embedded_images(os.path.join(ROOT, "assets/models/DamagedHelmet.gltf")):
            p = os.path.join(tmp, "helmet_%d.jpg" % i)
            open(p, "wb").write(data)
            jobs.append((p, False))
        for path, hdr in jobs:
            full = path if os.path.isabs(path) else os.path.join(ROOT, path)
            key = ("DamagedHelmet.gltf#image%s" % os.path.basename(path)[7]) if path.startswith(tmp) else path
            for flip in (False, True):
                w, h, c, raw, arr = run(full, hdr, flip)
                fx["%s%s" % (key, "|flip" if flip else "")] = {
                    "w": w, "h": h, "file_channels": c, "hdr": hdr, "sha256": hashlib.sha256(raw).hexdigest(),
                    "probes": [[y, x] + [float(v) if hdr else int(v) for v in arr[y, x]] for y, x in probes(h, w)]}
                print(key, flip, w, h, c)
    json.dump(fx, open(os.path.join(HERE, "stb_decode_fixtures.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
