"""GPU binning (gudni_b200/csrc/binning.cu, level 2 of the C ABI) against the hand-worked tile-tree fixtures of
tests/golden/tiletree_handworked.py — derived on paper from Raster/TileTree.hs:81-190, Raster/Job.hs:132-178 and
OpenCL/CallKernels.hs:244-255, not produced by any implementation."""
import numpy as np
import pytest

from gudni_b200.raster import setup_rasterizer
from gudni_b200.scene import SceneBuilder
from tiletree_cases import CASES, expected_tiles, spec_of

pytestmark = pytest.mark.gpu


def rectangles_scene(case):
    """Axis-aligned rectangles whose boxes are the fixture's (real geometry, so the frame can be rasterized)."""
    w, h = case["canvas"]
    b = SceneBuilder(w, h, (1.0, 1.0, 1.0, 1.0), name="handworked")
    colour = b.solid(0.2, 0.4, 0.6, 0.5)
    for (l, t, r, bt, _) in case["shapes"]:
        b.rectangle(colour, float(r - l), float(bt - t), [("translate", float(l), float(t))])
    scene = b.freeze()
    assert scene.culled == 0 and len(scene.entries) == len(case["shapes"])
    for i, (l, t, r, bt, strands) in enumerate(case["shapes"]):
        e = scene.entries[i]
        assert (e["left"], e["top"], e["right"], e["bottom"], e["num_strands"]) == (l, t, r, bt, strands)
    return scene


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_binning_matches_hand_derivation(name):
    case = CASES[name]
    scene = rectangles_scene(case)
    r = setup_rasterizer(spec=spec_of(case))
    try:
        r.debug_enable(True)
        img, stats = r.raster_scene(0, scene)
        tiles, shapes = r.debug_binned()
    finally:
        r.close()
    want, numbers = expected_tiles(case)
    assert len(tiles) == len(want)
    for field in ("left", "top", "right", "bottom", "h_depth", "v_depth", "shape_start", "shape_count"):
        assert np.array_equal(tiles[field], want[field]), (field, np.flatnonzero(tiles[field] != want[field])[:8])
    number_of = {int(g): i for i, g in enumerate(scene.entries["geo_start"])}
    got = np.asarray([number_of[int(g)] for g in shapes["geo_start"]], np.int64)
    assert np.array_equal(got, numbers)
    if "jobs" in case:   # column allocation as the swapped arguments give it (CallKernels.hs:254, Job.hs:144-178)
        j = case["jobs"]
        k = np.arange(len(tiles))
        assert np.array_equal(tiles["column_allocation"], j["column_step"] * (k % j["tiles_per_job"]))
    assert stats.n_tiles == len(want) and stats.n_shape_refs == len(numbers)
