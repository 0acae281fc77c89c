"""Render one of gudni_b200.scenes through the C ABI (level 2) and write it as a PPM — the headless output target.
usage: python tools/render_ppm.py <scene> [out.ppm]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gudni_b200 import headless, scenes  # noqa: E402
from gudni_b200.raster import setup_rasterizer  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "translucent_stack"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", which + ".ppm")
scene = getattr(scenes, which)()
r = setup_rasterizer()
img, stats = r.raster_scene(0, scene)
r.close()
os.makedirs(os.path.dirname(out), exist_ok=True)
headless.write_ppm(out, img)
print(out, img.shape, stats.as_dict())
