"""Shared helpers for the parity tests."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    """Switch arguments written "@GOLD/name" in the fixture are files next to it (quantization tables, scan scripts)."""
    with open(os.path.join(GOLD, "golden.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        c["switches"] = [os.path.join(GOLD, s[6:]) if s.startswith("@GOLD/") else s for s in c["switches"]]
    return cases


def case_id(c):
    im = c["image"] if isinstance(c["image"], str) else "synth%s%dx%d" % ("12b_" if len(c["image"]) > 3 else "", c["image"][1], c["image"][2])
    return im + ":" + "_".join(os.path.basename(s).lstrip("-") for s in c["switches"])


_img_cache = {}


def case_image(c):
    from oracle import oracle as O
    import mozjpeg_b200 as mj
    key = json.dumps(c["image"])
    if key not in _img_cache:
        if c["image"] == "testorig":
            w, h, nc, data = mj.read_ppm(open(os.path.join(GOLD, "testorig.ppm"), "rb").read())
            _img_cache[key] = np.frombuffer(data, dtype=np.uint8).reshape(h, w, nc)
        else:
            seed, w, h = c["image"][:3]
            if len(c["image"]) > 3:                      # [seed, w, h, 12]: 12-bit samples in uint16
                from mozjpeg_b200.synth import synth_image12
                _img_cache[key] = synth_image12(seed, w, h)
            else:
                _img_cache[key] = O.synth_image(seed, w, h)
    return _img_cache[key]


def md5(b):
    return hashlib.md5(b).hexdigest()


def device_supports(p):
    """Switch sets the device path does not cover yet are skipped, not faked."""
    import ctypes as C
    from mozjpeg_b200 import _abi as A
    return A.load().b200jpeg_validate(C.byref(p)) == 0
