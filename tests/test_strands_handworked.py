"""Strand building against fixtures worked out by hand from Raster/Strand.hs:74-178, Deknob.hs:61-108 and
ReorderTable.hs:44-110 (tests/golden/strands_handworked.py holds the derivations): the harness's restatement of the
Haskell (csrc/host/strand.hpp, through SceneBuilder.shape) and the per-shape logic of the level-3 kernels
(csrc/strand_build.cuh compiled for the host).  The fixtures are literals, not the output of any implementation."""
import numpy as np
import pytest

from gudni_b200.scene import SceneBuilder
from golden.strands_handworked import CASES, pairs_array
from test_strands_host import build, lib, parse_heap   # noqa: F401  (lib is a fixture)


def scene_of(case):
    b = SceneBuilder(32, 32)
    b.shape(b.solid(1, 0, 0, 1), [pairs_array(case)])
    return b.freeze()


def check(case, geometry, entries):
    assert len(entries) == 1
    e = entries[0]
    assert int(e["num_strands"]) == len(case["strands"])
    assert (e["left"], e["top"], e["right"], e["bottom"]) == tuple(np.float32(v) for v in case["box"])
    strands = parse_heap(geometry, entries)[0]
    for got, want in zip(strands, case["strands"]):
        assert np.array_equal(got, np.asarray(want, np.float32)), (got.tolist(), want)
    assert geometry.nbytes == sum(8 * (len(s) + 1) for s in case["strands"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_harness_matches_hand_derivation(name):
    scene = scene_of(CASES[name])
    check(CASES[name], scene.geometry, scene.entries)


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernel_logic_matches_hand_derivation(lib, name):   # noqa: F811
    geometry, entries, n = build(lib, scene_of(CASES[name]))
    check(CASES[name], geometry, entries)
    assert n == len(CASES[name]["strands"])
