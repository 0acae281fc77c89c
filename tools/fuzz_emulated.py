"""Long differential run, CPU only: the CUDA kernels under the host SIMT emulator (tests/native/raster_emu.cpp)
vs the oracle on seeded random small scenes — levels 1, 2 and 3, several RasterSpecs.
   python tools/fuzz_emulated.py <first case> <seconds> [--far]    prints one JSON line; MISMATCH <case> ... on a difference.
--far: scenes of shapes up to 1e30 pixels across (the case number is printed first: run it under `timeout`)."""
import ctypes
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from gudni_b200 import _build, scenes  # noqa: E402
from gudni_b200.formats import CSpec, RasterSpec  # noqa: E402
import test_kernels_emulated as T  # noqa: E402

SPECS = [RasterSpec(), RasterSpec(64, 64, 64, 512, 510, 127), RasterSpec(32, 32, 32, 256, 254, 127),
         RasterSpec(128, 128, 128, 1024, 1022, 127), RasterSpec(1024, 1024, 1024, 2853, 2851, 127)]


FAR = "--far" in sys.argv      # shapes up to 1e30 pixels across (scenes.far_shapes) instead of the ordinary mix


def main():
    if FAR:
        sys.argv.remove("--far")
    case = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    c = ctypes
    L = ctypes.CDLL(_build.build_raster_emu())
    vp, i32, i64, sz = c.c_void_p, c.c_int, c.c_int64, c.c_size_t
    L.raster_emu_frame.argtypes = [vp, sz, vp, vp, vp, vp, i32, i32, c.POINTER(CSpec), vp, i64, vp, vp, i32, i64, vp, vp, vp, vp]
    L.raster_emu_scene.argtypes = [vp, sz, vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, c.POINTER(CSpec), vp,
                                   vp, sz, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp]
    t0 = time.time()
    n, bad, skipped = 0, [], 0
    while time.time() - t0 < seconds:
        rng = np.random.default_rng(700000 + case)
        w, h = int(rng.integers(8, 160)), int(rng.integers(8, 120))
        kind = 3 if FAR else int(rng.integers(0, 3))
        if kind == 3:
            print("case", case, flush=True)      # a case that never returns is identified by the last line printed
            sc = scenes.far_shapes(int(rng.integers(1, 30)), w, h, 830000 + case)
        elif kind == 0:
            sc = scenes.mixed_bag(int(rng.integers(1, 120)), w, h, 800000 + case)
        elif kind == 1:
            sc = scenes.fuzzy_circles(int(rng.integers(1, 400)), w, h, float(rng.uniform(0.5, 5)), float(rng.uniform(5, 60)), 810000 + case)
        else:
            sc = scenes.random_rectangles(int(rng.integers(1, 150)), w, h, 820000 + case, max_size=float(rng.uniform(3, 100)))
        spec = SPECS[int(rng.integers(0, len(SPECS)))]
        level = int(rng.integers(1, 4))
        try:
            if level == 1:
                T.run(L, sc, spec)
            else:
                T.run_scene(L, sc, level, spec)
        except AssertionError as e:
            if "overflow_threads" in traceback.format_exc():
                skipped += 1
            else:
                bad.append(case)
                print("MISMATCH", case, level, spec, repr(e)[:200], flush=True)
        n += 1
        case += 1
    print(json.dumps({"cases": n, "skipped_overflow": skipped, "mismatches": bad, "seconds": time.time() - t0, "next_case": case}))


if __name__ == "__main__":
    main()
