import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer, DeviceScene
from oracle import oracle
which = sys.argv[1] if len(sys.argv) > 1 else "s4"
s = getattr(scenes, which)()
print("scene", s.name, s.n_shapes, len(s.geometry))
r = setup_rasterizer()
for i in range(3):
    t=time.time(); img, st = r.raster_scene(i, s); print("gpu l2 frame wall %.1f ms" % ((time.time()-t)*1e3), st.as_dict())
d = DeviceScene(r, s)
for i in range(5):
    r.frame_begin_device(d, i); r.raster_entries_device(d.entries, s.n_shapes); _, st = r.frame_end(want_image=False)
    print("resident: bin %.3f raster %.3f ms" % (st.ms_bin, st.ms_raster))
if "--check" in sys.argv:
    t=time.time(); ref = oracle.render(s, taps=False); print("oracle %.2f s on %d threads" % (time.time()-t, oracle.host_threads()))
    print("mismatch pixels", int((img != ref.image).sum()))
