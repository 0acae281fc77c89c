import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer, DeviceScene
a, b = int(sys.argv[1]), int(sys.argv[2])
s = scenes.s5()
r = setup_rasterizer()
ent = s.subset_rows(a, b)
d = DeviceScene(r, s, entries=ent)
for i in range(6):
    r.frame_begin_device(d, i); r.frame_strip(a, b); r.raster_entries_device(d.entries, d.n_entries); _, st = r.frame_end(want_image=False)
print("strip %d-%d: tiles %d bin %.3f raster %.3f ms" % (a, b, st.n_tiles, st.ms_bin, st.ms_raster))
if "--check" in sys.argv:
    img, st = r.frame_end(want_image=True) if False else (None, None)
