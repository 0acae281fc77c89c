"""A/B of library variants (gpurun_in/lib_*.so + the in-tree one) on one tile-row strip of S5 and on whole S5/S4:
device-resident raster ms per variant.  usage: python tools/ab_strip.py [row0 row1]"""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gudni_b200 import _build, raster, scenes
a, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4608, 6656)
s5, s4 = scenes.s5(), scenes.s4()
libs = [("tree", os.path.join(ROOT, "gudni_b200", "libgudni_b200.so"))]
libs += [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_in", "lib_*.so")))]
for name, path in libs:
    raster._lib = None
    _build.LIB_CUDA = path
    r = raster.setup_rasterizer()
    out = []
    for scene, rows in ((s5, (a, b)), (s5, None), (s4, None)):
        ent = scene.subset_rows(*rows) if rows else scene.entries
        d = raster.DeviceScene(r, scene, entries=ent)
        t = []
        for i in range(7):
            r.frame_begin_device(d, i)
            if rows:
                r.frame_strip(*rows)
            r.raster_entries_device(d.entries, d.n_entries)
            _, st = r.frame_end(want_image=False)
            if i >= 3:
                t.append(st.ms_raster + st.ms_bin)
        d.free()
        out.append(float(np.mean(t)))
    r.close()
    print(f"{name:24s} S5 strip {a}-{b}: {out[0]:7.3f} ms   S5: {out[1]:7.3f} ms   S4: {out[2]:7.3f} ms", flush=True)
