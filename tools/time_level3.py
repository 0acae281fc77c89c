"""Level 3 (outlines in, strands built on the GPU) on one scene, device-resident: per-stage times from the
library's own events.  Under `ncu --metrics gpu__time_duration.sum` this gives the launch list of the three
strand kernels beside the binning and raster kernels.   python tools/time_level3.py [scene] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200 import scenes  # noqa: E402
from gudni_b200.raster import DeviceScene, setup_rasterizer  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "s4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
s = getattr(scenes, which)()
r = setup_rasterizer()
d = DeviceScene(r, s)
d.put_outlines()
print("scene", s.name, "raw shapes", len(s.raw[0]), "outlines", len(s.raw[1]), "pairs", len(s.raw[2]), "transforms", len(s.raw[3]))
for i in range(frames):
    r.frame_begin_device_outlines(d, i)
    r.raster_outlines_device(d)
    _, st = r.frame_end(want_image=False)
    print("level 3 resident: strands %.3f  bin %.3f  raster %.3f ms" % (st.ms_strands, st.ms_bin, st.ms_raster))
