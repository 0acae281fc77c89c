"""How much the wide slice pass buys: a scene whose columns all hold a run of `n` thresholds that start together.
usage: python tools/time_replay_wide.py [n ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gudni_b200.raster import setup_rasterizer, DeviceScene
from gudni_b200.scene import SceneBuilder

def identical_shapes(n, width, height):
    b = SceneBuilder(width, height, (1.0, 1.0, 1.0, 1.0), name=f"identical-{n}")
    for i in range(n):
        b.rectangle(b.solid(0.1 + 0.8 * (i % 7) / 7.0, 0.5, 0.9 - 0.8 * (i % 5) / 5.0, 0.35), width * 0.8, height * 0.7,
                    [("translate", width * 0.1, height * 0.1), ("rotate", 0.03)])
    return b.freeze()

r = setup_rasterizer()
for n in [int(a) for a in sys.argv[1:]] or [12, 13, 40, 72, 73]:
    s = identical_shapes(n, 1920, 1080)
    d = DeviceScene(r, s)
    for i in range(4):
        r.frame_begin_device(d, i); r.raster_entries_device(d.entries, s.n_shapes); _, st = r.frame_end(want_image=False)
    print("run of %3d: raster %.3f ms, replayed threads %d" % (n, st.ms_raster, st.n_spilled_threads), flush=True)
    d.free()
