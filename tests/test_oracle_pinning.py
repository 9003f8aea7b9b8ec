"""CPU: pin the oracle (oracle/jpeg_oracle.c) to the reference.

* against the reference's own golden vector (testimages/testimgint.jpg ==
  `cjpeg -revert -dct int testorig.ppm`, md5 in CMakeLists.txt:1391) and the
  md5s recorded from the unmodified reference by tools/make_golden.py;
* live, byte for byte, against oracle/_ref when it is built (container and GPU
  box both carry it; skipped otherwise).
"""
import os

import numpy as np
import pytest

from common import case_id, case_image, golden_cases, md5

CASES = golden_cases()
SMALL = [c for c in CASES if c["image"] == "testorig" or c["image"][1] * c["image"][2] <= 200 * 136 or 65500 in c["image"][1:3]]


def test_reference_golden_md5_is_the_cmake_one():
    assert CASES[0]["switches"] == ["-revert", "-dct", "int"] and CASES[0]["image"] == "testorig"
    assert CASES[0]["md5"] == "9a68f56bc76e466aa7e52f415d0f4a5f"      # MD5_JPEG_420_ISLOW


@pytest.mark.parametrize("case", SMALL, ids=case_id)
def test_oracle_matches_recorded_reference(built, case):
    import mozjpeg_b200 as mj
    from oracle import oracle as O
    img = case_image(case)
    nc = 1 if img.ndim == 2 else img.shape[2]
    p = mj.params_from_switches(case["switches"], img.shape[1], img.shape[0], nc)
    out = O.oracle_encode(p, img).jpeg
    assert len(out) == case["size"] and md5(out) == case["md5"]


@pytest.mark.parametrize("size", [(640, 480), (1920, 1080)], ids=lambda s: "%dx%d" % s)
def test_oracle_large_cases(built, size):
    import mozjpeg_b200 as mj
    from oracle import oracle as O
    for c in CASES:
        if c["image"] != "testorig" and tuple(c["image"][1:]) == size and c["switches"][0] in ("-baseline", "-fastcrush") and "75" in c["switches"] and "2x2" in c["switches"]:
            img = case_image(c)
            p = mj.params_from_switches(c["switches"], img.shape[1], img.shape[0], 3)
            assert md5(O.oracle_encode(p, img).jpeg) == c["md5"]


def _fullsize_subset():
    """The full-size fixture's cases the restatement finishes in seconds (the 4K scan search takes it a minute)."""
    import json
    from common import GOLD
    path = os.path.join(GOLD, "fullsize_golden.json")
    cases = json.load(open(path))["cases"] if os.path.exists(path) else []
    keep = []
    for c in cases:
        sw = c["switches"]
        if c["image"][0] in (300, 17, 26) and sw[0] in ("-baseline", "-fastcrush", "-precision"):
            keep.append(c)
    return keep


@pytest.mark.parametrize("case", _fullsize_subset(), ids=case_id)
def test_oracle_full_size_recorded_reference(built, case):
    """BASELINE.json's configurations at their stated sizes (4K baseline + trellis, 4K progressive, 1080p q50/q90,
    12-bit 4:4:4 4K): the restatement reproduces the reference's md5."""
    import mozjpeg_b200 as mj
    from oracle import oracle as O
    img = case_image(case)
    p = mj.params_from_switches(case["switches"], img.shape[1], img.shape[0], 3)
    out = O.oracle_encode(p, img).jpeg
    assert len(out) == case["size"] and md5(out) == case["md5"]


def test_oracle_live_vs_reference_random_shapes(built):
    """Odd shapes x profiles, live against the compiled reference."""
    import mozjpeg_b200 as mj
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    sws = [["-baseline", "-quality", "70"], ["-fastcrush", "-quality", "80"], ["-revert", "-optimize"],
           ["-baseline", "-quality", "75", "-sample", "2x1"], ["-baseline", "-quality", "75", "-sample", "1x2"]]
    for _ in range(12):
        w, h = int(rng.integers(1, 97)), int(rng.integers(1, 97))
        img = O.synth_image(int(rng.integers(0, 1 << 30)), w, h)
        for sw in sws:
            p = mj.params_from_switches(sw, w, h)
            assert O.oracle_encode(p, img).jpeg == O.ref_encode(img, sw), (w, h, sw)


def test_stage_oracles_vs_reference_internals(built):
    """jpeg_fdct_islow of the reference library vs our restatement."""
    import ctypes as C
    from oracle import oracle as O
    if not O.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for _ in range(200):
        blk = rng.integers(-128, 160, 64).astype(np.int32)
        a = blk.copy(); b = blk.copy()
        O.orc().orc_fdct_islow(a.ctypes.data_as(C.POINTER(C.c_int)))
        O.ref().refshim_fdct_islow(b.ctypes.data_as(C.POINTER(C.c_int)))
        assert (a == b).all()


def test_random_switch_sets_live_against_reference_cjpeg(built):
    """tools/fuzz_vs_reference.py, a short run: random cjpeg switch sets (profiles, quality, sampling, restarts, DCT,
    smoothing, tuning presets, lambda, DC weight) on random small images - reference binary vs mirror + oracle."""
    import subprocess
    import sys
    from common import ROOT
    from oracle import oracle as O
    if not (O.ref_available() and os.path.exists(os.path.join(O.REF_DIR, "cjpeg"))):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_vs_reference.py"), "2024", "60"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-500:]
