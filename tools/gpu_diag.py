#!/usr/bin/env python3
"""Stage-by-stage comparison of the device path with the CPU oracle.
Usage: python tools/gpu_diag.py [W H [switches...]]   (needs a GPU)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200JPEG_KEEP_PLAIN", "1")
import numpy as np
import mozjpeg_b200 as mj
from oracle import oracle as O


def diag(w, h, sw, nimg=2, verbose=True):
    imgs = np.stack([O.synth_image(100 + s, w, h) for s in range(nimg)])
    p = mj.params_from_switches(sw, w, h)
    enc = mj.Encoder(0)
    ok = True
    try:
        out = enc.encode_batch(p, imgs)
    except Exception as e:
        print("ENCODE FAILED", (w, h), sw, e)
        return False
    for i in range(nimg):
        r = O.oracle_encode(p, imgs[i], want_debug=True)
        d = r.dbg
        same = out[i] == r.jpeg
        if verbose or not same:
            print(f"[{w}x{h} {' '.join(sw)}] image {i}: bytes {'OK' if same else 'MISMATCH'} dev={len(out[i])} orc={len(r.jpeg)}")
        if same:
            continue
        ok = False
        for ci in range(d["ncomp"]):
            for plane, key in ((1, "raw"), (2, "plain"), (0, "final")):
                dev = enc.debug_coefs(i, ci, plane)
                ref = d[key][ci]
                hib, wib = d["hib"][ci], d["wib"][ci]
                if plane == 1:
                    dev = dev[:hib, :wib]; ref = ref[:hib, :wib]
                bad = np.argwhere((dev != ref).any(axis=2))
                print(f"   comp {ci} {key:5s}: {len(bad)} differing blocks of {dev.shape[0]*dev.shape[1]}", end="")
                if len(bad):
                    r0, c0 = bad[0]
                    print(f"  first at (row {r0}, col {c0})\n      dev {dev[r0, c0].tolist()}\n      ref {ref[r0, c0].tolist()}")
                else:
                    print()
            if p.trellis_quant:
                for is_ac, key in ((False, "trellis_dc"), (True, "trellis_ac")):
                    tbl = p.comp_info[ci].ac_tbl_no if is_ac else p.comp_info[ci].dc_tbl_no
                    dv = enc.debug_huff(i, -1 - ci, is_ac, tbl)
                    print(f"   comp {ci} {key}: {'OK' if dv == d[key][ci] else 'MISMATCH'}")
                    if dv != d[key][ci]:
                        print("      dev", dv); print("      ref", d[key][ci])
        for si in range(d["nscans"]):
            for is_ac, key in ((False, "scan_dc"), (True, "scan_ac")):
                for t in range(2):
                    try:
                        dv = enc.debug_huff(i, si, is_ac, t)
                    except Exception as e:
                        dv = ("err", str(e))
                    rf = d[key][si][t]
                    if sum(rf[0]) and dv != rf:
                        print(f"   scan {si} {key}[{t}] MISMATCH\n      dev {dv}\n      ref {rf}")
        # first differing byte
        a, b = out[i], r.jpeg
        k = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), min(len(a), len(b)))
        print(f"   first differing byte at {k}: dev {a[k:k+16].hex()} ref {b[k:k+16].hex()}")
    enc.close()
    return ok


if __name__ == "__main__":
    if len(sys.argv) >= 3:
        w, h = int(sys.argv[1]), int(sys.argv[2]); sw = sys.argv[3:] or ["-baseline", "-quality", "75"]
        sys.exit(0 if diag(w, h, sw) else 1)
    allok = True
    for (w, h) in [(16, 16), (227, 149), (200, 136), (640, 480)]:
        for sw in (["-revert"], ["-revert", "-optimize"], ["-baseline", "-notrellis", "-quality", "75"], ["-baseline", "-quality", "75"],
                   ["-baseline", "-quality", "75", "-sample", "1x1"], ["-baseline", "-quality", "90", "-sample", "2x1"], ["-baseline", "-grayscale", "-quality", "60"]):
            allok &= diag(w, h, sw, verbose=False)
    print("ALL OK" if allok else "SOME MISMATCHES")
    sys.exit(0 if allok else 1)
