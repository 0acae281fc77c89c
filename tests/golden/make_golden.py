"""Generates the golden vectors under tests/golden/ from the REFERENCE'S OWN KERNELS.

`oracle/_ref/libgudni_ref.so` is src/Graphics/Gudni/OpenCL/Kernels.cl of the reference compiled for the
host (oracle/refbuild/build_ref.py; needs /root/reference, so this script runs in the build container,
not on the GPU box).  For a fixed list of scenes it records, per scene:

  sha256            of the BGRA8 image the reference's renderThresholds wrote
  thresholds        total qSlice.sLength after the reference's generateThresholds
  counts_sha256     of the per-thread threshold counts (int32, threadId order, -1 = inactive thread)
  bits_sha256       of the per-thread ShapeState.shapeBits

into reference_hashes.json, plus the images of the smallest scenes verbatim (reference_images.npz) so a
failing hash can be looked at.  The tile binning that feeds the kernels (Raster/TileTree.hs, Haskell —
not compilable here) is the restated one; the kernels see the same jobs either way.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gudni_b200 import scenes  # noqa: E402
from gudni_b200.formats import RasterSpec  # noqa: E402

SMALL_SPEC = RasterSpec(max_tile_size=64, threads_per_tile=64, max_tiles_per_call=64, max_thresholds=256,
                        max_strands_per_tile=254)

# name -> (scene factory, spec or None for the canonical one)
SCENES = {
    "tiny_square": (scenes.tiny_square, None),
    "medium_square": (scenes.medium_square, None),
    "full_rectangle": (scenes.full_rectangle, None),
    "stack_of_squares": (scenes.stack_of_squares, None),
    "open_square": (scenes.open_square, None),
    "concentric_squares2": (scenes.concentric_squares2, None),
    "concentric_squares3": (scenes.concentric_squares3, None),
    "six_point_rectangle": (scenes.six_point_rectangle, None),
    "hour_glass": (scenes.hour_glass, None),
    "translucent_stack": (scenes.translucent_stack, None),
    "square_100_0.4": (lambda: scenes.square(100, 0.4), None),
    "square_512_0.625": (lambda: scenes.square(512, 0.625), None),
    "random_shapes_7": (lambda: scenes.random_rectangles(120, 200, 160, 7), None),
    "fuzzy_circles_small": (lambda: scenes.fuzzy_circles(400, 256, 192, 5, 40, 0x601D), None),
    "fuzzy_circles_small_tiles": (lambda: scenes.fuzzy_circles(300, 200, 120, 4, 30, 0x601E), SMALL_SPEC),
    "picture_scene": (lambda: scenes.picture_scene(320, 300, flowers_size=(350, 200)), None),
    "paragraph_small": (lambda: scenes.s2(480, 200, lines=4), None),
    "thin_rectangles": (lambda: scenes.thin_rectangles(40), None),
}
VERBATIM = ("tiny_square", "medium_square", "open_square", "hour_glass", "translucent_stack")

# BASELINE.json's configs at their stated sizes (reference_hashes_fullsize.json, `--fullsize`): minutes of CPU
# time, so they are kept apart from the quick set above.  The -m gpu tests hold the CUDA path to them at
# levels 1 and 2 (tests/test_gpu_fullsize.py).
FULLSIZE = {
    "s2_1920x1080": (lambda: scenes.s2(1920, 1080), None),
    "s3_3840x2160": (lambda: scenes.s3(3840, 2160), None),
    "s4b_3840x2160": (scenes.s4b, None),
    "s4_3840x2160": (scenes.s4, None),
    "s5_16384": (scenes.s5, None),
    "s5b_16384": (scenes.s5b, None),
}


def digest(result):
    counts = np.concatenate(result.n_thresholds).astype("<i4")
    bits = np.concatenate(result.shape_bits).astype("<i4")
    return {"sha256": hashlib.sha256(result.image.astype("<u4").tobytes()).hexdigest(),
            "thresholds": int(result.total_thresholds),
            "counts_sha256": hashlib.sha256(counts.tobytes()).hexdigest(),
            "bits_sha256": hashlib.sha256(bits.tobytes()).hexdigest()}


def render(name, reference):
    from oracle import oracle
    make, spec = SCENES[name] if name in SCENES else FULLSIZE[name]
    kw = {} if spec is None else {"spec": spec}
    return oracle.render(make(), reference=reference, **kw)


if __name__ == "__main__":
    from oracle import oracle
    if oracle.reference_lib() is None:
        sys.exit("needs the reference tree (/root/reference) to compile its kernels")
    if "--fullsize" in sys.argv:
        out = {}
        for name in FULLSIZE:
            r = render(name, reference=True)
            assert r.overflow_threads == 0, name
            out[name] = digest(r)
            out[name]["canvas"] = [int(r.image.shape[1]), int(r.image.shape[0])]
            # the restated oracle on the same jobs: the two must agree before anything is written
            mine = digest(render(name, reference=False))
            assert all(mine[k] == out[name][k] for k in mine), (name, mine, out[name])
            print(name, json.dumps(out[name]), flush=True)
            del r
        with open(os.path.join(HERE, "reference_hashes_fullsize.json"), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
        sys.exit(0)
    out, images = {}, {}
    for name in SCENES:
        r = render(name, reference=True)
        assert r.overflow_threads == 0, name
        out[name] = digest(r)
        if name in VERBATIM:
            images[name] = r.image.astype("<u4")
    with open(os.path.join(HERE, "reference_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "reference_images.npz"), **images)
    print(json.dumps(out, indent=1, sort_keys=True))
