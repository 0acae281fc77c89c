"""Long differential run, CPU only: the level-3 kernels' per-shape logic compiled for the host
(tests/native/strand_check.cpp) vs the harness's restatement of Raster/Strand.hs on seeded mixed bags.
   python tools/fuzz_strands.py <seconds>"""
import sys, time, json, ctypes, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200 import scenes, _build
from gudni_b200.formats import ENTRY_DTYPE
L = ctypes.CDLL(_build.build_strand_check()); c=ctypes
L.strand_check_build.restype = c.c_int64
L.strand_check_build.argtypes = [c.c_void_p, c.c_int, c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int, c.c_void_p, c.c_size_t, c.POINTER(c.c_size_t), c.c_void_p, c.POINTER(c.c_int64)]
def build(scene):
    shapes, outlines, pairs, transforms = (np.ascontiguousarray(a) for a in scene.raw)
    nbytes, nstr = c.c_size_t(), c.c_int64()
    ptr = lambda a: a.ctypes.data if a.size else None
    kept = L.strand_check_build(ptr(shapes), len(shapes), ptr(outlines), ptr(pairs), ptr(transforms), scene.width, scene.height, None, 0, c.byref(nbytes), None, c.byref(nstr))
    g = np.zeros(nbytes.value, np.uint8); e = np.zeros(kept, ENTRY_DTYPE)
    L.strand_check_build(ptr(shapes), len(shapes), ptr(outlines), ptr(pairs), ptr(transforms), scene.width, scene.height, g.ctypes.data, g.nbytes, c.byref(nbytes), e.ctypes.data, c.byref(nstr))
    return g,e
t0=time.time(); n=0; bad=[]
case=0
while time.time()-t0 < float(sys.argv[1]):
    rng=np.random.default_rng(500000+case)
    sc=scenes.mixed_bag(int(rng.integers(1,400)),int(rng.integers(8,900)),int(rng.integers(8,700)),600000+case)
    g,e=build(sc)
    ok = len(e)==sc.n_shapes and e.tobytes()==sc.entries.tobytes() and g.tobytes()==sc.geometry.tobytes()
    if not ok: bad.append(case); print("MISMATCH",case,flush=True)
    n+=1; case+=1
print(json.dumps({"cases":n,"mismatches":bad,"seconds":time.time()-t0}))
