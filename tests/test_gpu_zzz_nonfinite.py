"""Coordinates at the far end of float32.  (1) Boxes beyond int32 root tiles are binned like any other.
(2) Geometry with a point at +-infinity.  The reference's curve bisection (Kernels.cl:1226-1258) does not terminate
on it — its own kernels hang — so the library refuses the frame instead: strand_bounds_kernel raises a flag,
tile_order_kernel empties the shape lists of the launch, the raster kernels paint background only, and
gudni_b200_frame_end returns GUDNI_ERR_ARGUMENT.  The context stays usable.

The check runs in a child process under a time limit: if the guard ever failed the symptom would be a kernel that
never returns, and that must cost this test, not the session.  (Sorted last among the GPU tests for the same reason.)
The CPU suite holds the same kernels to the same behaviour under the emulator
(tests/test_kernels_emulated.py::test_infinite_coordinate_is_refused_not_rasterized)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys
import numpy as np
sys.path.insert(0, {root!r})
from gudni_b200 import scenes
from gudni_b200.raster import GudniError, setup_rasterizer
from oracle import oracle          # the checker: job layout of the finite scene, reference image afterwards

r = setup_rasterizer()
for level in (1, 2):
    for coord in (0, 1):
        for value in (np.inf, -np.inf):
            scene = scenes.medium_square()
            jobs = oracle.build_raster_jobs(scene)
            scene.geometry = scene.geometry.copy()
            scene.geometry.view(np.float32)[2 * 2 + coord] = value     # unit 0 is the strand header
            try:
                if level == 1:
                    r.queue_raster_jobs(0, scene, jobs)
                else:
                    r.raster_scene(0, scene)
            except GudniError as e:
                assert e.code == -1, e.code                              # GUDNI_ERR_ARGUMENT
                assert "infinity" in str(e), str(e)
            else:
                raise SystemExit(f"level {{level}}: frame with {{value}} accepted")
            # the context is still good, and the next (finite) frame is bit-exact
            scene = scenes.medium_square()
            ref = oracle.render(scene, taps=False)
            img, stats = r.raster_scene(1, scene) if level == 2 else r.queue_raster_jobs(1, scene, ref.jobs)
            assert np.array_equal(img, ref.image)
            assert stats.n_thresholds == ref.total_thresholds
r.close()
print("REFUSED-AND-RECOVERED")
"""


def test_boxes_beyond_int32_are_binned(rasterizer):
    """Level 2 and 3 on a scene whose boxes, in root tiles, overflow an int32: tiles, shape lists, counts and pixels
    against the oracle (the binning clamps its candidate range in float)."""
    from gudni_b200 import scenes
    from parity import level2_parity
    for scene in [scenes.huge_boxes()] + [scenes.far_shapes(12, 150, 110, 0xFA50 + seed) for seed in range(3)]:
        img, stats, ref = level2_parity(rasterizer, scene)
        img3, stats3 = rasterizer.raster_outlines(0, scene)
        assert (img3 == ref.image).all() and stats3.n_thresholds == ref.total_thresholds, scene.name


def test_infinite_coordinate_is_refused():
    try:
        done = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT)], cwd=ROOT, capture_output=True, text=True,
                              timeout=240)
    except subprocess.TimeoutExpired:
        pytest.fail("the child did not return: a raster kernel is looping on the infinite coordinate")
    assert done.returncode == 0, done.stdout[-2000:] + done.stderr[-4000:]
    assert "REFUSED-AND-RECOVERED" in done.stdout
