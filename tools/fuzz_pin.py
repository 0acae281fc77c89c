"""Long differential run, CPU only: restated oracle vs the reference's own kernels compiled for the host
(oracle/_ref) on seeded random scenes (mixed bags, circles, rectangles; four RasterSpecs; canvases 8..500 px).
   python tools/fuzz_pin.py <first case> <last case> <seconds> [--far]    prints one JSON line; MISMATCH <case> on a difference.
--far: scenes.far_shapes (sizes up to 1e30 pixels) instead of the ordinary mix; the case number is printed first."""
import sys, time, json, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200 import scenes
from gudni_b200.formats import RasterSpec
from oracle import oracle
SPECS=[None, RasterSpec(64,64,64,512,510,127), RasterSpec(32,32,32,256,254,127), RasterSpec(128,128,128,1024,1022,127)]
FAR = '--far' in sys.argv
if FAR: sys.argv.remove('--far')
t0=time.time(); n=0; bad=[]; skipped=0
case=int(sys.argv[1]); end=int(sys.argv[2])
while case<end and time.time()-t0 < float(sys.argv[3]):
    rng=np.random.default_rng(100000+case)
    w,h=int(rng.integers(8,500)),int(rng.integers(8,400))
    kind=int(rng.integers(0,3))
    if FAR:
        print('case', case, flush=True); sc=scenes.far_shapes(int(rng.integers(1,40)),w,h,500000+case)
    elif kind==0: sc=scenes.mixed_bag(int(rng.integers(1,250)),w,h,200000+case)
    elif kind==1: sc=scenes.fuzzy_circles(int(rng.integers(1,600)),w,h,float(rng.uniform(0.5,6)),float(rng.uniform(6,80)),300000+case)
    else: sc=scenes.random_rectangles(int(rng.integers(1,300)),w,h,400000+case,max_size=float(rng.uniform(3,150)))
    spec=SPECS[int(rng.integers(0,4))]
    kw={} if spec is None else {"spec":spec}
    jobs=oracle.build_raster_jobs(sc,**kw)
    a=oracle.raster_jobs(sc,jobs,**kw)
    if a.overflow_threads: skipped+=1; case+=1; continue
    b=oracle.raster_jobs(sc,jobs,reference=True,**kw)
    eq = all(np.array_equal(x,y) for x,y in zip(a.n_thresholds,b.n_thresholds)) and all(np.array_equal(x,y) for x,y in zip(a.shape_bits,b.shape_bits)) and np.array_equal(a.image,b.image)
    if not eq: bad.append(case); print("MISMATCH", case, flush=True)
    n+=1; case+=1
print(json.dumps({"cases":n,"skipped_overflow":skipped,"mismatches":bad,"seconds":time.time()-t0,"last_case":case}))
