"""Scenes that live on the replay path (queues over the on-chip capacity, tiles with more shapes than stack
bits): device time per frame, spilled threads."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer, DeviceScene
r = setup_rasterizer()
cases = {
    "300 thresholds per column (1 shape, 150 thin rectangles, 256x256)": scenes.thin_rectangles(150, width=256, height=256, spacing=1.5, thickness=0.7, one_shape=True),
    "6000 circles on 128x128 (8-px tiles over MAXSHAPE)": scenes.fuzzy_circles(6000, 128, 128, 5, 40, 0xB175),
    "20000 circles on 512x512": scenes.fuzzy_circles(20000, 512, 512, 5, 40, 0xB176),
}
for name, s in cases.items():
    d = DeviceScene(r, s)
    for i in range(3):
        r.frame_begin_device(d, i); r.raster_entries_device(d.entries, s.n_shapes); _, st = r.frame_end(want_image=False)
    print(f"{name}: raster {st.ms_raster:.2f} ms, tiles {st.n_tiles}, spilled threads {st.n_spilled_threads}, overflow {st.n_overflow_threads}")
    d.free()
