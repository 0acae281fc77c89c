"""Generates tests/golden/oracle_hashes.json: sha256 of the oracle's BGRA image and its threshold
total for a fixed list of small scenes.  The reference itself cannot be run (no GHC / OpenCL in the
image), so these are regression pins of the restated algorithm, not reference outputs.
Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gudni_b200 import scenes  # noqa: E402

SCENES = {
    "tiny_square": scenes.tiny_square,
    "open_square": scenes.open_square,
    "concentric_squares3": scenes.concentric_squares3,
    "hour_glass": scenes.hour_glass,
    "square_100_0.4": lambda: scenes.square(100, 0.4),
    "square_512_0.625": lambda: scenes.square(512, 0.625),
    "random_shapes_7": lambda: scenes.random_rectangles(120, 200, 160, 7),
    "fuzzy_circles_small": lambda: scenes.fuzzy_circles(400, 256, 192, 5, 40, 0x601D),
    "picture_scene": lambda: scenes.picture_scene(320, 300, flowers_size=(350, 200)),
    "paragraph_small": lambda: scenes.s2(480, 200, lines=4),
}

if __name__ == "__main__":
    from oracle import oracle
    out = {}
    for name, make in SCENES.items():
        r = oracle.render(make())
        out[name] = {"sha256": hashlib.sha256(r.image.tobytes()).hexdigest(), "thresholds": int(r.total_thresholds)}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_hashes.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))
