"""Extension parameters the reference exposes only through jpeg_c_set_*_param (no cjpeg switch): the trellis split into
two AC bands (use_scans_in_trellis / trellis_freq_split, jcmaster.c:451-467).  md5s recorded from the unmodified
reference (tools/make_golden.py -> tests/golden/ext_golden.json)."""
import hashlib
import json
import os

import pytest

from common import GOLD

EXT = json.load(open(os.path.join(GOLD, "ext_golden.json")))["cases"]


def _id(c):
    return "%dx%d:%s:%s" % (c["width"], c["height"], "_".join(s.lstrip("-") for s in c["switches"]), "_".join("%s%d" % (k.split("_")[-1], v) for k, v in c["ext"].items()))


def _inputs(c):
    import mozjpeg_b200 as mj
    from oracle import oracle as O
    p = mj.params_from_switches(c["switches"], c["width"], c["height"], 3)
    for k, v in c["ext"].items():
        setattr(p, k, v)
    return p, O.synth_image(c["seed"], c["width"], c["height"])


@pytest.mark.parametrize("c", EXT, ids=_id)
def test_oracle_matches_recorded_reference(built, c):
    from oracle import oracle as O
    p, im = _inputs(c)
    out = O.oracle_encode(p, im).jpeg
    assert len(out) == c["size"] and hashlib.md5(out).hexdigest() == c["md5"]


@pytest.mark.gpu
@pytest.mark.parametrize("c", EXT, ids=_id)
def test_device_matches_recorded_reference(encoder, c):
    from common import device_supports
    p, im = _inputs(c)
    if not device_supports(p):
        pytest.skip("parameter set not on the device path (B200JPEG_ERR_UNSUPPORTED): skipped, not faked")
    out = encoder.encode_batch(p, im[None])[0]
    assert len(out) == c["size"] and hashlib.md5(out).hexdigest() == c["md5"]
