#!/usr/bin/env python3
"""Test infrastructure, like tools/fuzz_vs_reference.py (needs oracle/_ref): random raw-data (jpeg_write_raw_data) inputs and random extension-parameter sets (optional trellis modes, DC weight):
reference vs oracle.
usage: fuzz_raw_ext.py seed cases"""
import sys, random; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import oracle as O
import mozjpeg_b200 as mj
from mozjpeg_b200.synth import synth_planes
rng=random.Random(int(sys.argv[1]))
bad=0; tot=0
for it in range(int(sys.argv[2])):
    w=rng.choice([1,8,17,33,64,100,131]); h=rng.choice([1,8,23,40,64])
    # raw data
    sw=[rng.choice(["-baseline","-fastcrush","-revert",""])]; sw=[x for x in sw if x]+["-quality",str(rng.choice([30,75,90]))]
    if rng.random()<0.6: sw+=["-sample",rng.choice(["1x1","2x1","1x2","2x2","4x2","2x2,1x1,2x2"])]
    gray = rng.random()<0.2
    if gray: sw+=["-grayscale"]
    try:
        p=mj.params_from_switches(sw,w,h,1 if gray else 3)
        planes=synth_planes(p,rng.randrange(1<<20))
        a=O.ref_encode_raw(planes,w,h,sw); b=O.oracle_encode_raw(p,planes); tot+=1
        if a!=b: bad+=1; print("MISMATCH-RAW",sw,(w,h),len(a),len(b))
    except Exception as ex: print("EXC-RAW",sw,(w,h),str(ex)[:120]); bad+=1
    # ext params
    im=O.synth_image(rng.randrange(1<<20),w,h)
    sw=[rng.choice(["-baseline","-fastcrush",""])]; sw=[x for x in sw if x]+["-quality",str(rng.choice([20,50,75,90]))]
    if rng.random()<0.4: sw+=["-sample",rng.choice(["1x1","2x1","2x2"])]
    if rng.random()<0.2: sw+=["-restart","1"]
    if rng.random()<0.2: sw+=["-dct",rng.choice(["float","fast"])]
    if rng.random()<0.2: sw+=["-trellis-dc-ver-weight","1.5"]
    ext={}
    if rng.random()<0.5: ext["use_scans_in_trellis"]=1
    if rng.random()<0.4: ext["trellis_freq_split"]=rng.choice([1,2,8,30,62])
    if rng.random()<0.4: ext["trellis_num_loops"]=rng.choice([1,2,3])
    if rng.random()<0.3: ext["trellis_q_opt"]=1
    if rng.random()<0.3: ext["trellis_eob_opt"]=1
    try:
        a=O.ref_encode(im,sw,ext)
        p=mj.params_from_switches(sw,w,h,3)
        for k,v in ext.items(): setattr(p,k,v)
        b=O.oracle_encode(p,im).jpeg; tot+=1
        if a!=b: bad+=1; print("MISMATCH-EXT",sw,ext,(w,h),len(a),len(b))
    except Exception as ex: print("EXC-EXT",sw,ext,(w,h),str(ex)[:120]); bad+=1
print("seed", sys.argv[1], "bad", bad, "compared", tot)
