"""A/B harness for kernel variants, one process, one GPU call:

    tools/variants.sh base "" fewer_sections "-DGUDNI_SECTIONS_PER_ROUND=5" ...     # builds gpurun_in/lib_<name>.so
    gpurun -- python tools/ab_variants.py [scene] [frames]                           # on the GPU box

Every library found in gpurun_in/ (and the in-tree one as "tree") renders the scene device-resident, is
timed from the library's own events (bin + raster ms, best and mean of `frames`), and has its image hashed
against the golden vector of the reference's kernels under OpenCL on a B200
(tests/golden/opencl_b200_hashes.json) — a variant that is fast and wrong is marked WRONG, not ranked.
Results go to stdout and gpurun_out/ab_variants.json."""
import glob
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from gudni_b200 import _build, raster, scenes  # noqa: E402


def run(path, scene, frames, golden):
    raster._lib = None                      # make load_library() bind the variant
    _build.LIB_CUDA = path
    r = raster.setup_rasterizer()
    try:
        img, stats = r.raster_scene(0, scene)
        digest = hashlib.sha256(img.astype("<u4").tobytes()).hexdigest()
        d = raster.DeviceScene(r, scene)
        times = []
        for i in range(3 + frames):
            r.frame_begin_device(d, i)
            r.raster_entries_device(d.entries, scene.n_shapes)
            _, st = r.frame_end(want_image=False)
            if i >= 3:
                times.append((st.ms_bin, st.ms_raster))
        d.free()
    finally:
        r.close()
    t = np.asarray(times)
    return {"lib": os.path.relpath(path, ROOT), "bin_ms": float(t[:, 0].mean()), "raster_ms_mean": float(t[:, 1].mean()),
            "raster_ms_best": float(t[:, 1].min()), "thresholds": int(stats.n_thresholds),
            "spilled": int(stats.n_spilled_threads), "correct": (digest == golden["sha256"]) if golden else None}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "s4"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    scene = getattr(scenes, which)()
    with open(os.path.join(ROOT, "tests", "golden", "opencl_b200_hashes.json")) as f:
        golden = json.load(f)["strict"].get(which)
    libs = [("tree", os.path.join(ROOT, "gudni_b200", "libgudni_b200.so"))]
    libs += [(os.path.basename(p)[4:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_in", "lib_*.so")))]
    results = {}
    for name, path in libs:
        # each variant gets its own dlopen handle: ctypes caches by path, the paths differ
        try:
            results[name] = run(path, scene, frames, golden)
        except Exception as e:  # noqa: BLE001
            results[name] = {"error": repr(e)[:300]}
        r = results[name]
        print(f"{name:28s}", "ERROR " + r["error"] if "error" in r else
              f"raster {r['raster_ms_mean']:7.3f} ms (best {r['raster_ms_best']:7.3f})  bin {r['bin_ms']:.3f}  "
              f"{'ok' if r['correct'] else ('WRONG' if r['correct'] is False else 'unchecked')}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ab_variants.json"), "w") as f:
        json.dump({"scene": scene.name, "frames": frames, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
