"""compute-sanitizer scenario (GPU box): small scenes through every raster path — dense tiles, pictures, the wide slice pass,
the lane-private replay (tiles over MAXSHAPE, queues over the on-chip capacity), level 3, a frame stored into a host bitmap as
two batches on two streams — each checked against the oracle.
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
    compute-sanitizer --tool synccheck python tools/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer
from gudni_b200.scene import SceneBuilder
from oracle import oracle

def identical_shapes(n, width=48, height=40):
    b = SceneBuilder(width, height, (1.0, 1.0, 1.0, 1.0), name=f"identical-{n}")
    for i in range(n):
        b.rectangle(b.solid(0.1 + 0.8 * (i % 7) / 7.0, 0.5, 0.9 - 0.8 * (i % 5) / 5.0, 0.35), 20.3, 25.7, [("translate", 9.2, 6.1), ("rotate", 0.03)])
    return b.freeze()

r = setup_rasterizer()
cases = [scenes.fuzzy_circles(120, 160, 120, 4, 40, 6), scenes.picture_scene(160, 150, flowers_size=(175, 100)), identical_shapes(30),
         identical_shapes(80), scenes.fuzzy_circles(700, 64, 64, 5, 40, 77), scenes.thin_rectangles(140, width=64, height=224, spacing=1.5, thickness=0.7, one_shape=True)]
for k, s in enumerate(cases):
    ref = oracle.render(s, taps=False).image
    img, st = r.raster_scene(k, s)
    assert np.array_equal(img, ref), s.name
    print("ok", s.name, "replayed", st.n_spilled_threads, flush=True)
s = cases[0]
ref = oracle.render(s, taps=False).image
img, st = r.raster_outlines(10, s)
assert np.array_equal(img, ref)
host = np.zeros((s.height, s.width), np.uint32)
r.host_register(host)
r.frame_target_host(host)
r.raster_scene(11, s, out=host)
assert np.array_equal(host, ref)
r.frame_target_host(None)
r.host_unregister(host)
print("sanitizer scenario complete")
