"""GUDNI_BATCHES sweep on the GPU box: S4, S5 and one eighth of S5 (the per-rank load of the 8-GPU run), device-resident,
timed from the library's own events; every setting's image is compared with the first one's.
usage: python tools/ab_batches.py [batches ...]      -> stdout + gpurun_out/ab_batches.json"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from gudni_b200 import scenes  # noqa: E402
from gudni_b200.raster import setup_rasterizer, DeviceScene  # noqa: E402

settings = sys.argv[1:] or ["1", "2", "2o", "3o", "4o"]      # "3o": three ordered batches (GUDNI_BATCH_ORDERED); "NAME=VALUE[,NAME=VALUE]": any environment
frames = 8
cases = [("s4", scenes.s4(), None), ("s5", scenes.s5(), None), ("s5_strip", None, (4608, 6656))]
out = {}
for name, scene, strip in cases:
    if scene is None:
        scene = cases[1][1]
    ref = None
    for b in settings:
        if "=" in b:
            for kv in b.split(","):
                k, v = kv.split("=")
                if v == "": os.environ.pop(k, None)
                else: os.environ[k] = v
        else:
            os.environ["GUDNI_BATCHES"] = b.rstrip("o")
            os.environ["GUDNI_BATCH_ORDERED"] = "1" if b.endswith("o") else "0"
        r = setup_rasterizer()
        ent = scene.subset_rows(*strip) if strip else None
        d = DeviceScene(r, scene, entries=ent) if strip else DeviceScene(r, scene)
        n = d.n_entries if strip else scene.n_shapes
        times = []
        img = None
        for i in range(3 + frames):
            r.frame_begin_device(d, i)
            if strip:
                r.frame_strip(*strip)
            r.raster_entries_device(d.entries, n)
            last = i == 2 + frames
            img, st = r.frame_end(want_image=last)
            if i >= 3:
                times.append(st.ms_raster)
        digest = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
        ref = ref or digest
        d.free()
        r.close()
        t = np.asarray(times)
        out[f"{name}/{b}"] = {"raster_ms_mean": float(t.mean()), "raster_ms_best": float(t.min()), "same_image": digest == ref,
                              "spilled": int(st.n_spilled_threads)}
        print(f"{name:9s} batches {b}: raster {t.mean():7.3f} ms (best {t.min():7.3f})  {'same image' if digest == ref else 'DIFFERENT IMAGE'}  spilled {st.n_spilled_threads}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_batches.json"), "w"), indent=1)
