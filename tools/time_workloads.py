import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer, DeviceScene
r = setup_rasterizer()
for name in sys.argv[1:]:
    t = time.time(); s = getattr(scenes, name)(); tb = time.time() - t
    d = DeviceScene(r, s)
    times = []
    for i in range(4):
        r.frame_begin_device(d, i); r.raster_entries_device(d.entries, s.n_shapes); _, st = r.frame_end(want_image=False)
        times.append((st.ms_bin, st.ms_raster))
    print(f"{name}: shapes {s.n_shapes} geo {len(s.geometry)/1e6:.1f}MB build {tb:.1f}s tiles {st.n_tiles} refs {st.n_shape_refs} thr {st.n_thresholds} spilled {st.n_spilled_threads} overflow {st.n_overflow_threads} bin {times[-1][0]:.3f} raster {times[-1][1]:.3f} ms (first {times[0][1]:.1f})", flush=True)
    d.free()
