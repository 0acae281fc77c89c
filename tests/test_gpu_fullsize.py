"""BASELINE.json's configs at their stated sizes, through the C ABI on the B200.

Two independent checks per config:
  * against tests/golden/reference_hashes_fullsize.json — the reference's own Kernels.cl compiled for the
    host (tests/golden/make_golden.py --fullsize): image SHA-256, threshold total, hashes of the per-thread
    threshold counts and shape-bit counts at level 1, and the image hash again at level 2 (tiles binned on
    the GPU);
  * live against the oracle at level 2: tile boxes / depths / order / shape lists, per-thread counts, pixels.
S5 / S5b are 16384^2 canvases (1 GiB of pixels): the oracle needs tens of seconds on the box's cores.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from gudni_b200.formats import CANONICAL_SPEC
from oracle import oracle

from golden.make_golden import FULLSIZE, digest
from parity import level2_parity

pytestmark = pytest.mark.gpu

_PATH = os.path.join(os.path.dirname(__file__), "golden", "reference_hashes_fullsize.json")
GOLDEN = json.load(open(_PATH)) if os.path.exists(_PATH) else {}


class _Result:
    def __init__(self, image, counts, bits, total):
        self.image, self.n_thresholds, self.shape_bits, self.total_thresholds = image, [counts], [bits], total


@pytest.mark.parametrize("name", sorted(FULLSIZE))
def test_fullsize_golden_levels_1_and_2(rasterizer, name):
    assert name in GOLDEN, "run tests/golden/make_golden.py --fullsize where /root/reference exists"
    scene = FULLSIZE[name][0]()
    assert [scene.width, scene.height] == GOLDEN[name]["canvas"]
    jobs = oracle.build_raster_jobs(scene, CANONICAL_SPEC)
    rasterizer.debug_enable(True)
    img, stats = rasterizer.queue_raster_jobs(0, scene, jobs)
    counts, bits = rasterizer.debug_thread_counts()
    rasterizer.debug_enable(False)
    assert stats.n_overflow_threads == 0
    got = digest(_Result(img, counts, bits, stats.n_thresholds))
    want = {k: GOLDEN[name][k] for k in got}
    assert got == want
    del img, counts, bits
    img2, stats2 = rasterizer.raster_scene(1, scene)
    assert stats2.n_thresholds == GOLDEN[name]["thresholds"]
    assert hashlib.sha256(img2.astype("<u4").tobytes()).hexdigest() == GOLDEN[name]["sha256"]


@pytest.mark.parametrize("name", ["s3_3840x2160", "s5_16384", "s5b_16384"])
def test_fullsize_level2_against_the_oracle(rasterizer, name):
    scene = FULLSIZE[name][0]()
    level2_parity(rasterizer, scene)
