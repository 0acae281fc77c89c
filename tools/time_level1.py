"""Level 1 of the ABI on a big scene: tiles binned by the oracle's tile tree (as the Haskell caller would),
one gudni_b200_raster_job per job of <= G tiles.  Prints wall and device times per frame."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer
from oracle import oracle
which = sys.argv[1] if len(sys.argv) > 1 else "s4"
s = getattr(scenes, which)()
jobs = oracle.build_raster_jobs(s)
print("scene", s.name, "jobs", len(jobs), "tiles", sum(len(j.tiles) for j in jobs))
r = setup_rasterizer()
for i in range(4):
    t = time.time(); img, st = r.queue_raster_jobs(i, s, jobs); wall = (time.time() - t) * 1e3
    print("level 1 frame: wall %.1f ms, device raster %.2f ms, upload %.2f ms" % (wall, st.ms_raster, st.ms_upload))
img2, st2 = r.raster_scene(9, s)
print("level 2 frame: device raster %.2f ms bin %.2f ms; images equal: %s" % (st2.ms_raster, st2.ms_bin, bool((img == img2).all())))
