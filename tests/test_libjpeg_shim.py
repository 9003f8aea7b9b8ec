"""The drop-in boundary exercised the way an application would: the reference's
own, unmodified `cjpeg` binary (oracle/_ref/cjpeg, built by oracle/Makefile)
runs with integration/_build/libjpeg_b200shim.so in front of the reference's
libjpeg, so jpeg_start_compress / jpeg_write_scanlines / jpeg_finish_compress
go to the device path.  B200_SHIM_REQUIRE=1 makes a fall-through to the
reference's code an error, so a pass here cannot come from the CPU encoder."""
import hashlib
import os
import subprocess

import pytest

from common import GOLD, ROOT, golden_cases

SHIM = os.path.join(ROOT, "integration", "_build", "libjpeg_b200shim.so")
CJPEG = os.path.join(ROOT, "oracle", "_ref", "cjpeg")
PPM = os.path.join(GOLD, "testorig.ppm")

need_files = pytest.mark.skipif(not (os.path.exists(SHIM) and os.path.exists(CJPEG)),
                                reason="shim / reference cjpeg not built (needs /root/reference at build time)")


def _run(switches, env_extra, tmp_path):
    out = tmp_path / "o.jpg"
    env = dict(os.environ, LD_PRELOAD=SHIM, B200_SHIM_VERBOSE="1", **env_extra)
    r = subprocess.run([CJPEG, *switches, "-outfile", str(out), PPM], env=env, capture_output=True, text=True, timeout=300)
    return r, (out.read_bytes() if out.exists() else b"")


def _golden(switches):
    for c in golden_cases():
        if c["image"] == "testorig" and c["switches"] == switches:
            return c
    return None


DEVICE_SETS = [["-baseline", "-quality", "75"], ["-quality", "75", "-fastcrush"], ["-revert", "-dct", "int"],
               ["-baseline", "-notrellis", "-quality", "75"], ["-revert", "-optimize", "-grayscale"],
               ["-quality", "75"],                      # the library default: 64-candidate scan search
               ["-baseline", "-quality", "75", "-smooth", "30"]]


@need_files
@pytest.mark.gpu
@pytest.mark.parametrize("sw", DEVICE_SETS, ids=lambda s: "_".join(x.lstrip("-") for x in s))
def test_reference_cjpeg_runs_on_the_device(sw, tmp_path):
    r, data = _run(sw, {"B200_SHIM_REQUIRE": "1"}, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "device path" in r.stderr, r.stderr
    plain = subprocess.run([CJPEG, *sw, PPM], capture_output=True, timeout=300)      # the reference itself, no shim
    assert plain.returncode == 0
    assert data == plain.stdout
    g = _golden(sw)
    if g:
        assert hashlib.md5(data).hexdigest() == g["md5"]


@need_files
@pytest.mark.gpu
def test_unsupported_parameters_fall_through_to_the_reference(tmp_path):
    """Arithmetic coding is not on the device path: the shim must hand the image to the
    reference's own implementation, and say so."""
    r, data = _run(["-quality", "75", "-arithmetic"], {}, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "reference path" in r.stderr
    plain = subprocess.run([CJPEG, "-quality", "75", "-arithmetic", PPM], capture_output=True, timeout=300)
    assert data == plain.stdout


@need_files
def test_shim_without_gpu_is_transparent(tmp_path):
    """No CUDA device (the CPU test box) or MOZ_B200_FORCE_CPU: the reference encodes, bytes unchanged;
    with B200_SHIM_REQUIRE=1 the same situation is an error, never a silent CPU result."""
    r, data = _run(["-baseline", "-quality", "75"], {"MOZ_B200_FORCE_CPU": "1"}, tmp_path)
    assert r.returncode == 0, r.stderr
    assert "reference path" in r.stderr
    plain = subprocess.run([CJPEG, "-baseline", "-quality", "75", PPM], capture_output=True, timeout=300)
    assert data == plain.stdout
    r2, _ = _run(["-baseline", "-quality", "75"], {"MOZ_B200_FORCE_CPU": "1", "B200_SHIM_REQUIRE": "1"}, tmp_path)
    assert r2.returncode != 0
