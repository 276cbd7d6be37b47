"""Extracts the reference's PMJ02BN and blue-noise sampler tables into binary files the build can load.

The reference keeps them as C arrays (src/lib/vengine/math/PMJSequences.cpp: 16 x 16384 x 2 floats, from
github.com/Andrew-Helmer/pmj-cpp "pmj02bn 16k_samples_01..16"; src/lib/vengine/math/BlueNoise.cpp: 48 x 128 x 128 floats, from
github.com/MomentsInGraphics/BlueNoise) and uploads them verbatim as two storage buffers (vulkan/resources/VulkanRandom.cpp:40-72).
This script parses the literals (no reference code is copied, only the numbers) and writes
    assets/tables/pmj02bn.f32       little-endian float32 [16][16384][2]        (2 MiB)
    assets/tables/bluenoise.u16     little-endian uint16  [48][128][128], value = u16 / 65536  (1.5 MiB; every entry of the
                                    reference table is a multiple of 2^-16, which the script verifies)
    assets/tables/tables.json       shapes + SHA-256 of the float32 images of both tables
usage: python tools/extract_sampler_tables.py [/root/reference]"""
import hashlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
math_dir = os.path.join(ref, "src", "lib", "vengine", "math")
num = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?=F?\s*[,}])")


def literals(path, start_token):
    text = open(path).read()
    body = text[text.index(start_token):]
    body = body[body.index("{"):]
    return np.array([float(m.group(0)) for m in num.finditer(body)], np.float64)


pmj = literals(os.path.join(math_dir, "PMJSequences.cpp"), "pmj02bnSequences")
assert pmj.size == 16 * 16384 * 2, pmj.size
pmj32 = pmj.astype(np.float32)
assert np.all((pmj32 >= 0) & (pmj32 < 1))
blue = literals(os.path.join(math_dir, "BlueNoise.cpp"), "bluNoiseTextures")
assert blue.size == 48 * 128 * 128, blue.size
q = blue * 65536.0
assert np.all(q == np.round(q)) and q.min() >= 0 and q.max() < 65536, "blue-noise entries are not multiples of 2^-16"
blue16 = q.astype(np.uint16)
blue32 = (blue16.astype(np.float32) / np.float32(65536.0))
assert np.array_equal(blue32, blue.astype(np.float32))
out = os.path.join(ROOT, "assets", "tables")
os.makedirs(out, exist_ok=True)
pmj32.astype("<f4").tofile(os.path.join(out, "pmj02bn.f32"))
blue16.astype("<u2").tofile(os.path.join(out, "bluenoise.u16"))
meta = {"pmj02bn": {"shape": [16, 16384, 2], "dtype": "float32", "sha256_f32": hashlib.sha256(pmj32.astype("<f4").tobytes()).hexdigest()},
        "bluenoise": {"shape": [48, 128, 128], "stored": "uint16 / 65536", "sha256_f32": hashlib.sha256(blue32.astype("<f4").tobytes()).hexdigest()},
        "source": "kwstanths/vviewer src/lib/vengine/math/{PMJSequences,BlueNoise}.cpp (data from Andrew-Helmer/pmj-cpp and MomentsInGraphics/BlueNoise)"}
json.dump(meta, open(os.path.join(out, "tables.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(meta, indent=1))
