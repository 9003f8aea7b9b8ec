#!/usr/bin/env python3
"""Test infrastructure, like tools/fuzz_vs_reference.py (needs oracle/_ref): random 12-bit encodes (refshim) and random cjpeg -> jpegtran chains: reference vs mirrors + oracle.  Reference decoder
warnings on scan-search + row-restart sources (a reference quirk) show up as EXC-T lines and are not ours.
usage: fuzz_transcode_12bit.py seed cases"""
import sys, random; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import ctypes as C
from oracle import oracle as O
import mozjpeg_b200 as mj
from mozjpeg_b200 import _abi as A, jpegtran as T
from mozjpeg_b200.synth import synth_image12
rng=random.Random(int(sys.argv[1]))
bad=0; tot=0
for it in range(int(sys.argv[2])):
    w=rng.choice([1,8,17,33,64,100]); h=rng.choice([1,8,23,40,64])
    # 12-bit encode
    im=synth_image12(rng.randrange(1<<20),w,h)
    sw=["-precision","12","-notrellis","-noovershoot","-quality",str(rng.choice([20,60,75,90,100]))]
    sw+=[rng.choice(["-baseline","-fastcrush","-revert","-progressive"])] if rng.random()<0.9 else []
    if rng.random()<0.5: sw+=["-sample",rng.choice(["1x1","2x1","1x2","2x2"])]
    if rng.random()<0.2: sw+=["-grayscale"]
    if rng.random()<0.3: sw+=["-restart",rng.choice(["1","2B"])]
    try:
        a=O.ref_encode(im,sw)
        p=mj.params_from_switches(sw,w,h,3)
        b=O.oracle_encode(p,im).jpeg
        tot+=1
        if a!=b: bad+=1; print("MISMATCH12",sw,(w,h),len(a),len(b))
    except Exception as ex:
        print("EXC12",sw,(w,h),str(ex)[:100]); bad+=1
    # transcode
    im8=O.synth_image(rng.randrange(1<<20),w,h)
    esw=[rng.choice(["-revert","-baseline","-fastcrush",""])]
    esw=[x for x in esw if x]+["-quality",str(rng.choice([10,50,75,92]))]
    if rng.random()<0.5: esw+=["-sample",rng.choice(["1x1","2x1","2x2","3x2","2x2,1x1,2x2"])]
    if rng.random()<0.2: esw+=["-grayscale"]
    if rng.random()<0.2: esw+=["-restart","1"]
    tsw=[]
    if rng.random()<0.3: tsw+=["-revert"]
    if rng.random()<0.3: tsw+=["-optimize"]
    if rng.random()<0.3: tsw+=["-progressive"]
    if rng.random()<0.2: tsw+=["-fastcrush"]
    if rng.random()<0.2: tsw+=["-restart",rng.choice(["1","3B"])]
    try:
        src=O.ref_encode(im8,esw)
        co=O.ref_read_coefs(src)["coefs"]
        pp,ps=T.params_for_transcode(T.parse_header(src),tsw)
        b=O.oracle_encode_coefs(pp,co)
        if ps and pp.compress_profile==A.PROFILE_MAX_COMPRESSION and len(src)<len(b): b=src
        a=O.ref_jpegtran(src,tsw)
        tot+=1
        if a!=b: bad+=1; print("MISMATCH-T",esw,tsw,(w,h),len(a),len(b))
    except Exception as ex:
        print("EXC-T",esw,tsw,(w,h),str(ex)[:100]); bad+=1
print("seed", sys.argv[1], "bad", bad, "compared", tot)
