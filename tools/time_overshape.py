"""One scene whose tiles stay over MAXSHAPE at the 8-pixel floor (every column-thread takes the lane-private replay).
usage: python tools/time_overshape.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer, DeviceScene
r = setup_rasterizer()
s = scenes.fuzzy_circles(6000, 128, 128, 5, 40, 0xB175)
d = DeviceScene(r, s)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    r.frame_begin_device(d, i); r.raster_entries_device(d.entries, s.n_shapes); _, st = r.frame_end(want_image=False)
    print(f"raster {st.ms_raster:.2f} ms, tiles {st.n_tiles}, replayed threads {st.n_spilled_threads}")
