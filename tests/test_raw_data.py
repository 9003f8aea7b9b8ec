"""Raw-data input (jpeg_write_raw_data / tj3CompressFromYUV semantics): component planes in, same files out.
The md5s in tests/golden/raw_golden.json were recorded from the unmodified reference (tools/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from common import GOLD

RAW = json.load(open(os.path.join(GOLD, "raw_golden.json")))["cases"]


def _id(c):
    return "%dx%d:%s" % (c["width"], c["height"], "_".join(s.lstrip("-") for s in c["switches"]))


def _inputs(c):
    import mozjpeg_b200 as mj
    from mozjpeg_b200.synth import synth_planes
    p = mj.params_from_switches(c["switches"], c["width"], c["height"], 1 if "-grayscale" in c["switches"] else 3)
    return p, synth_planes(p, c["seed"])


@pytest.mark.parametrize("c", RAW, ids=_id)
def test_oracle_raw_matches_recorded_reference(built, c):
    from oracle import oracle as O
    p, planes = _inputs(c)
    out = O.oracle_encode_raw(p, planes)
    assert len(out) == c["size"] and hashlib.md5(out).hexdigest() == c["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("c", RAW, ids=_id)
def test_device_raw_matches_recorded_reference(encoder, c):
    p, planes = _inputs(c)
    out = encoder.encode_batch_raw(p, [a[None] for a in planes])[0]
    assert len(out) == c["size"] and hashlib.md5(out).hexdigest() == c["md5"]


@pytest.mark.gpu
def test_device_raw_batch(encoder):
    """Several images per call, padded plane pitch."""
    import mozjpeg_b200 as mj
    from mozjpeg_b200.synth import synth_planes
    from oracle import oracle as O
    sw = ["-baseline", "-quality", "75"]
    p = mj.params_from_switches(sw, 120, 72, 3)
    per = [synth_planes(p, 50 + i) for i in range(3)]
    stacked = []
    for ci in range(3):
        h, w = per[0][ci].shape
        buf = np.zeros((3, h + 3, w + 24), dtype=np.uint8)
        for i in range(3):
            buf[i, :h, :w] = per[i][ci]
        stacked.append(buf[:, :h, :w])                    # non-contiguous views are made contiguous by the binding
    out = encoder.encode_batch_raw(p, stacked)
    for i in range(3):
        assert out[i] == O.oracle_encode_raw(p, per[i])
