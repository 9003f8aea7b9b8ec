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


JPEGTRAN = os.path.join(ROOT, "oracle", "_ref", "jpegtran")
need_tran = pytest.mark.skipif(not (os.path.exists(SHIM) and os.path.exists(CJPEG) and os.path.exists(JPEGTRAN)),
                               reason="shim / reference cjpeg / jpegtran not built (needs /root/reference at build time)")


def _tran(switches, src, env_extra):
    env = dict(os.environ, LD_PRELOAD=SHIM, B200_SHIM_VERBOSE="1", **env_extra)
    return subprocess.run([JPEGTRAN, *switches], input=src, env=env, capture_output=True, timeout=300)


@need_tran
@pytest.mark.gpu
@pytest.mark.parametrize("tsw", [[], ["-progressive"], ["-revert", "-optimize"], ["-fastcrush", "-restart", "2"], ["-revert"],
                                 ["-rotate", "90"], ["-flip", "horizontal", "-progressive"], ["-grayscale", "-optimize"], ["-crop", "100x80+16+16"]],
                         ids=lambda s: "_".join(x.lstrip("-") for x in s) or "default")
@pytest.mark.parametrize("esw", [["-revert"], ["-quality", "85", "-sample", "1x1"]], ids=lambda s: "_".join(x.lstrip("-") for x in s))
def test_reference_jpegtran_encodes_on_the_device(esw, tsw):
    """jpeg_read_coefficients (and the lossless transforms, which fill the output arrays only after
    jpeg_write_coefficients) stay the reference's; jpeg_write_coefficients + jpeg_finish_compress run on the GPU."""
    src = subprocess.run([CJPEG, *esw, PPM], capture_output=True, timeout=300).stdout
    r = _tran(tsw, src, {"B200_SHIM_REQUIRE": "1"})
    assert r.returncode == 0, r.stderr
    assert b"device path (coefficients" in r.stderr, r.stderr
    plain = subprocess.run([JPEGTRAN, *tsw], input=src, capture_output=True, timeout=300)
    assert plain.returncode == 0 and r.stdout == plain.stdout


@need_tran
@pytest.mark.gpu
def test_application_markers_are_spliced_in_behind_the_file_header(tmp_path):
    """jpeg_write_icc_profile (jpeg_write_m_header / jpeg_write_m_byte) on the pixel path, jcopy_markers_execute
    (jpeg_write_marker) on the coefficient path."""
    icc = tmp_path / "fake.icc"
    icc.write_bytes(bytes(range(256)) * 300)                     # two APP2 chunks' worth is not needed; one 76 KB profile = 2 chunks
    sw = ["-quality", "75", "-icc", str(icc)]
    r, data = _run(sw, {"B200_SHIM_REQUIRE": "1"}, tmp_path)
    assert r.returncode == 0, r.stderr
    plain = subprocess.run([CJPEG, *sw, PPM], capture_output=True, timeout=300)
    assert plain.returncode == 0 and data == plain.stdout and b"ICC_PROFILE" in data
    t = _tran(["-copy", "all", "-progressive"], data, {"B200_SHIM_REQUIRE": "1"})
    assert t.returncode == 0, t.stderr
    plain_t = subprocess.run([JPEGTRAN, "-copy", "all", "-progressive"], input=data, capture_output=True, timeout=300)
    assert plain_t.returncode == 0 and t.stdout == plain_t.stdout and b"ICC_PROFILE" in t.stdout


TJBENCH = os.path.join(ROOT, "oracle", "_ref", "tjbench")
need_tj = pytest.mark.skipif(not (os.path.exists(SHIM) and os.path.exists(TJBENCH)),
                             reason="shim / reference tjbench not built (needs /root/reference at build time)")


@need_tj
@pytest.mark.gpu
@pytest.mark.parametrize("opts", [["-subsamp", "420"], ["-subsamp", "444", "-optimize"], ["-subsamp", "422", "-progressive"],
                                  ["-subsamp", "gray"], ["-subsamp", "420", "-restart", "1"],
                                  ["-subsamp", "420", "-yuv"], ["-subsamp", "422", "-yuv", "-optimize"]], ids=lambda s: "_".join(x.lstrip("-") for x in s))
@pytest.mark.parametrize("fmt", [None, "-rgb", "-rgbx", "-bgrx", "-xbgr", "-xrgb"], ids=lambda f: (f or "-bgr").lstrip("-"))
def test_turbojpeg_api_runs_on_the_device(opts, fmt, tmp_path):
    """The TurboJPEG API (tj3Compress8, turbojpeg.c:1268-1340) sits on the libjpeg API exactly as in the reference
    (turbojpeg-mp.c:104-125); with the reference's turbojpeg.c linked dynamically against libjpeg (oracle/Makefile) the
    shim is underneath it, and the reference's own tjbench writes the same files as without the shim, for every pixel
    format it offers (its default is TJPF_BGR -> JCS_EXT_BGR, turbojpeg.c:330-397).  -yuv goes through
    tj3EncodeYUV8 (CPU colour conversion, the reference's) + tj3CompressFromYUVPlanes8 -> jpeg_write_raw_data."""
    import shutil
    outs = []
    for name, env_extra in (("dev", {"LD_PRELOAD": SHIM, "B200_SHIM_REQUIRE": "1", "B200_SHIM_VERBOSE": "1"}), ("ref", {})):
        d = tmp_path / name
        d.mkdir()
        shutil.copyfile(PPM, d / "img.ppm")
        r = subprocess.run([TJBENCH, "img.ppm", "75", *([fmt] if fmt else []), "-quiet", "-benchtime", "0.01", "-warmup", "0", "-componly", *opts],
                           cwd=d, env=dict(os.environ, **env_extra), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
        if name == "dev":
            assert "device path" in r.stderr, r.stderr
        jpgs = sorted(f for f in os.listdir(d) if f.endswith(".jpg"))
        assert len(jpgs) == 1, os.listdir(d)
        outs.append((d / jpgs[0]).read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 100


@need_files
@pytest.mark.gpu
@pytest.mark.parametrize("sw", [["-precision", "12", "-quality", "75", "-notrellis", "-noovershoot", "-baseline"],
                                ["-precision", "12", "-quality", "85", "-notrellis", "-noovershoot", "-fastcrush", "-sample", "1x1"]],
                         ids=lambda s: "_".join(x.lstrip("-") for x in s))
def test_reference_cjpeg_12bit_runs_on_the_device(sw, tmp_path):
    """12-bit samples reach the library through jpeg12_write_scanlines (rows of J12SAMPLE = short)."""
    from mozjpeg_b200.synth import synth_image12
    im = synth_image12(5, 200, 136)
    ppm = tmp_path / "in12.ppm"
    ppm.write_bytes(b"P6\n200 136\n4095\n" + im.astype(">u2").tobytes())
    out = tmp_path / "o.jpg"
    env = dict(os.environ, LD_PRELOAD=SHIM, B200_SHIM_VERBOSE="1", B200_SHIM_REQUIRE="1")
    r = subprocess.run([CJPEG, *sw, "-outfile", str(out), str(ppm)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "device path" in r.stderr, r.stderr
    plain = subprocess.run([CJPEG, *sw, str(ppm)], capture_output=True, timeout=300)
    assert plain.returncode == 0 and out.read_bytes() == plain.stdout


# ---------------------------------------------------------------------------
# The standalone drop-in: integration/_build/libjpeg.so.62 is a complete libjpeg (API v6.2) -- the reference's own
# objects plus the shim, the 13 taken-over entry points of the reference renamed inside (integration/Makefile) -- that an
# application links or loads like the stock library; cjpeg_b200 is the reference's cjpeg front end linked against it.
# ---------------------------------------------------------------------------
STD_LIB = os.path.join(ROOT, "integration", "_build", "libjpeg.so.62")
STD_CJPEG = os.path.join(ROOT, "integration", "_build", "cjpeg_b200")
need_std = pytest.mark.skipif(not (os.path.exists(STD_LIB) and os.path.exists(STD_CJPEG) and os.path.exists(CJPEG)),
                              reason="standalone libjpeg.so.62 not built (needs /root/reference at build time)")


@need_std
def test_standalone_libjpeg_exports_the_v62_api():
    """SURVEY 8(b) "must export" list, with the reference's symbol versions (libjpeg.map.in)."""
    out = subprocess.check_output(["nm", "-D", "--defined-only", STD_LIB], text=True)
    have = {}
    for ln in out.splitlines():
        f = ln.split()
        if len(f) == 3 and f[1] == "T":
            name, _, ver = f[2].partition("@@")
            have[name] = ver
    want = ["jpeg_std_error", "jpeg_CreateCompress", "jpeg_destroy_compress", "jpeg_abort_compress", "jpeg_stdio_dest", "jpeg_mem_dest",
            "jpeg_set_defaults", "jpeg_set_colorspace", "jpeg_default_colorspace", "jpeg_set_quality", "jpeg_set_linear_quality",
            "jpeg_add_quant_table", "jpeg_quality_scaling", "jpeg_float_quality_scaling", "jpeg_simple_progression", "jpeg_suppress_tables",
            "jpeg_alloc_quant_table", "jpeg_alloc_huff_table", "jpeg_start_compress", "jpeg_write_scanlines", "jpeg12_write_scanlines",
            "jpeg_write_raw_data", "jpeg_finish_compress", "jpeg_write_marker", "jpeg_write_m_header", "jpeg_write_m_byte", "jpeg_write_tables",
            "jpeg_write_coefficients", "jpeg_c_bool_param_supported", "jpeg_c_set_bool_param", "jpeg_c_get_bool_param", "jpeg_c_int_param_supported",
            "jpeg_c_set_int_param", "jpeg_c_get_int_param", "jpeg_c_float_param_supported", "jpeg_c_set_float_param", "jpeg_c_get_float_param",
            "jpeg_CreateDecompress", "jpeg_read_header", "jpeg_read_coefficients"]
    missing = [w for w in want if w not in have]
    assert not missing, missing
    assert have["jpeg_start_compress"] == "LIBJPEG_6.2" and have["jpeg_mem_dest"] == "LIBJPEGTURBO_6.2"
    assert not [n for n in have if n.startswith("b200ref_")]              # the renamed reference entry points stay internal
    soname = subprocess.check_output(["readelf", "-d", STD_LIB], text=True)
    assert "libjpeg.so.62" in soname


@need_std
def test_standalone_libjpeg_without_a_device_is_the_reference():
    """No LD_PRELOAD, no GPU: the library hands the object to the reference code it carries (same bytes as the reference)."""
    env = dict(os.environ, MOZ_B200_FORCE_CPU="1")
    sw = ["-quality", "75", "-fastcrush"]
    a = subprocess.run([STD_CJPEG, *sw, PPM], env=env, capture_output=True, timeout=300)
    b = subprocess.run([CJPEG, *sw, PPM], capture_output=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0 and a.stdout == b.stdout and len(a.stdout) > 1000


@need_std
@pytest.mark.gpu
@pytest.mark.parametrize("sw", DEVICE_SETS[:4] + [["-quality", "75"]], ids=lambda s: "_".join(x.lstrip("-") for x in s))
def test_standalone_libjpeg_encodes_on_the_device(sw):
    env = dict(os.environ, B200_SHIM_REQUIRE="1", B200_SHIM_VERBOSE="1")
    a = subprocess.run([STD_CJPEG, *sw, PPM], env=env, capture_output=True, timeout=300)
    assert a.returncode == 0, a.stderr
    assert b"device path" in a.stderr, a.stderr
    b = subprocess.run([CJPEG, *sw, PPM], capture_output=True, timeout=300)
    assert b.returncode == 0 and a.stdout == b.stdout
