"""Render a scene a few times with a given build of the library (for `ncu` launch lists of a variant).
usage: python tools/time_lib.py path/to/lib.so [scene] [frames]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gudni_b200 import _build, raster, scenes  # noqa: E402
_build.LIB_CUDA = os.path.abspath(sys.argv[1])
raster._lib = None
scene = getattr(scenes, sys.argv[2] if len(sys.argv) > 2 else "s4")()
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
r = raster.setup_rasterizer()
d = raster.DeviceScene(r, scene)
for i in range(frames):
    r.frame_begin_device(d, i)
    r.raster_entries_device(d.entries, scene.n_shapes)
    _, st = r.frame_end(want_image=False)
    print("frame %d: bin %.3f raster %.3f ms" % (i, st.ms_bin, st.ms_raster))
