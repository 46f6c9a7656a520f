#!/usr/bin/env python
"""Measured FP64 peak of the GPU (pfx_fp64_peak: independent DFMA chains on every SM, CUDA events): the
denominator of the FP64 rooflines in bench.py.  Writes profiles/fp64_peak.json when asked to (--write)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from plateflex_b200 import _lib  # noqa: E402

t = max(_lib.fp64_peak(0) for _ in range(3))
out = {"tflops": t, "how": "pfx_fp64_peak: 8 independent DFMA chains per thread, 8 CTAs of 256 threads per SM, best of 4, CUDA events",
       "nominal_tflops": 37.0}
print(json.dumps(out))
if "--write" in sys.argv:
    json.dump(out, open(os.path.join(ROOT, "profiles", "fp64_peak.json"), "w"), indent=1)
