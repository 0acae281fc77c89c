"""Differential run on the GPU box: seeded random scenes (mixed bags with pictures and knobs, circles, rectangles; awkward
canvas sizes) through levels 2 and 3 of the C ABI against the oracle, bit-exact; every third scene with the frame stored into
a page-locked host bitmap (two batches on two streams).
    python tools/fuzz_gpu.py <first case> <seconds>       prints one JSON line; MISMATCH <case> on a difference."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from gudni_b200 import scenes
from gudni_b200.raster import setup_rasterizer
from oracle import oracle

first, seconds = int(sys.argv[1]), float(sys.argv[2])
r = setup_rasterizer()
t0 = time.time()
case, bad = first, []
while time.time() - t0 < seconds:
    rng = np.random.default_rng(case)
    w, h = int(rng.integers(20, 900)), int(rng.integers(20, 700))
    kind = case % 4
    if kind == 0: s = scenes.mixed_bag(int(rng.integers(1, 400)), w, h, 90000 + case)
    elif kind == 1: s = scenes.fuzzy_circles(int(rng.integers(5, 1500)), w, h, float(rng.uniform(1, 8)), float(rng.uniform(9, 120)), case)
    elif kind == 2: s = scenes.random_rectangles(int(rng.integers(3, 400)), w, h, case, max_size=float(rng.uniform(6, 200)))
    else: s = scenes.fuzzy_circles(int(rng.integers(200, 3000)), int(rng.integers(20, 200)), int(rng.integers(20, 200)), 5, 40, case)   # dense: wide runs, replay
    ref = oracle.render(s, taps=False)
    if ref.overflow_threads:
        case += 1
        continue
    host = None
    if case % 3 == 0:
        host = np.zeros((s.height, s.width), np.uint32)
        r.host_register(host)
        r.frame_target_host(host)
    img, st = r.raster_scene(case, s, out=host)
    ok = np.array_equal(img if host is None else host, ref.image)
    if host is not None:
        r.frame_target_host(None)
        r.host_unregister(host)
    if ok and kind != 3 and case % 2 == 0:
        img3, _ = r.raster_outlines(case, s)
        ok = np.array_equal(img3, ref.image)
    if not ok:
        bad.append(case)
        print("MISMATCH", case, s.name, flush=True)
    case += 1
print(json.dumps({"cases": case - first, "mismatches": bad, "seconds": time.time() - t0, "next_case": case}))
