"""Reads the texture unit's sRGB -> linear table through the product's parity hook (ptc_srgb_table) and writes it as a C header for
the oracle.  Run on a B200: python tools/read_srgb_table.py gpurun_out/srgb_table.h ; then copy to oracle/srgb_table.h."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vviewer_b200 import capi  # noqa: E402

ctx = capi.Context(capi.load_cuda())
t = ctx.srgb_table()
ctx.close()
c = np.arange(256, dtype=np.float64) / 255.0
analytic = np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)
rel = np.abs(t - analytic) / np.maximum(analytic, 1e-6)
with open(sys.argv[1], "w") as f:
    f.write("/* sRGB code -> linear value as NVIDIA texture units return it for VK_FORMAT_R8G8B8A8_SRGB / cudaTextureDesc::sRGB texels\n"
            " * (hardware-defined: SURVEY 8c(v)).  Read back on a B200 through ptc_srgb_table by tools/read_srgb_table.py; differs from the\n"
            " * analytic curve (c <= 0.04045 ? c / 12.92 : pow((c + 0.055) / 1.055, 2.4)) by up to %.2e relative (median %.2e).\n"
            " * tests/test_gpu_parity.py::test_srgb_table_is_the_hardware_s keeps it equal to what the device does. */\n" % (rel[1:].max(), np.median(rel[1:])))
    f.write("static const float kSrgbToLinear[256] = {\n")
    for i in range(0, 256, 4):
        f.write("    " + " ".join(("%.9g" % float(v) + ("" if any(ch in "%.9g" % float(v) for ch in ".e") else ".0") + "f,") for v in t[i:i + 4]) + "\n")
    f.write("};\n")
print("max rel diff to analytic %.3e, median %.3e; t[0]=%g t[255]=%g" % (rel[1:].max(), np.median(rel[1:]), t[0], t[255]))
