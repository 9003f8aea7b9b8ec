#!/usr/bin/env python3
"""Test infrastructure, like tools/fuzz_vs_reference.py (needs oracle/_ref): random VALID progressive scan scripts (band splits, successive approximation up to Al 3, DC refinement) through
cjpeg -scans: reference binary vs mirror + validate + oracle.
usage: fuzz_scan_scripts.py seed cases"""
import sys, random, tempfile, os; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import ctypes as C
from oracle import oracle as O
import mozjpeg_b200 as mj
from mozjpeg_b200 import _abi as A
rng=random.Random(int(sys.argv[1]))
bad=0; tot=0
for it in range(int(sys.argv[2])):
    gray = rng.random()<0.25
    nc = 1 if gray else 3
    scans=[]; later=[]
    al_dc=rng.choice([0,0,1,2])
    if nc==3 and rng.random()<0.5: scans.append(("0 1 2",0,0,0,al_dc))
    else:
        for c in range(nc): scans.append((str(c),0,0,0,al_dc))
    for k in range(al_dc,0,-1):
        if nc==3 and rng.random()<0.5: later.append(("0 1 2",0,0,k,k-1))
        else:
            for c in range(nc): later.append((str(c),0,0,k,k-1))
    for c in range(nc):
        cuts=sorted(rng.sample(range(2,63),rng.choice([0,1,2])))
        bands=[]; s0=1
        for cu in cuts: bands.append((s0,cu)); s0=cu+1
        bands.append((s0,63))
        for (ss,se) in bands:
            al=rng.choice([0,0,1,2,3])
            scans.append((str(c),ss,se,0,al))
            for k in range(al,0,-1): later.append((str(c),ss,se,k,k-1))
    rng.shuffle(later)
    # refinements of the same (comp, band) must stay in order: stable re-sort by decreasing Ah within key
    keyed={}
    for s in later: keyed.setdefault((s[0],s[1],s[2]),[]).append(s)
    for k in keyed: keyed[k].sort(key=lambda s:-s[3])
    out=[]; seen={}
    for s in later:
        k=(s[0],s[1],s[2]); i=seen.get(k,0); out.append(keyed[k][i]); seen[k]=i+1
    script=scans+out
    txt="".join("%s: %d %d %d %d;\n"%s for s in script)
    f=tempfile.NamedTemporaryFile("w",suffix=".txt",delete=False); f.write(txt); f.close()
    w=rng.choice([8,17,33,64,100]); h=rng.choice([8,23,40,64])
    im=O.synth_image(rng.randrange(1<<20),w,h)
    sw=["-scans",f.name,"-quality",str(rng.choice([30,75,92]))]
    if gray: sw+=["-grayscale"]
    if rng.random()<0.3: sw+=["-sample",rng.choice(["1x1","2x1","2x2"])]
    if rng.random()<0.2: sw+=["-restart",rng.choice(["1","3B"])]
    if rng.random()<0.2: sw=["-revert"]+sw
    try:
        a=O._ref_cjpeg_pixels(im,sw)
    except Exception as ex:
        p=mj.params_from_switches(sw,w,h,3); rc=A.load().b200jpeg_validate(C.byref(p))
        if rc==0: print("WE ACCEPT, ref rejects:",str(ex).strip()[-60:],txt.replace("\n"," ")); bad+=1
        os.unlink(f.name); continue
    p=mj.params_from_switches(sw,w,h,3)
    rc=A.load().b200jpeg_validate(C.byref(p))
    if rc: print("WE REJECT:",A.load().b200jpeg_last_error(),txt.replace("\n"," ")); bad+=1; os.unlink(f.name); continue
    b=O.oracle_encode(p,im).jpeg; tot+=1
    if a!=b: bad+=1; print("MISMATCH",sw[2:],(w,h),len(a),len(b),txt.replace("\n"," "))
    os.unlink(f.name)
print("seed", sys.argv[1], "bad", bad, "compared", tot)
