"""The oracle's tile tree and job packing (oracle/tiletree_oracle.cpp) against fixtures worked out by hand from
Raster/TileTree.hs:81-190, Raster/Job.hs:132-178 and OpenCL/CallKernels.hs:244-255 (tests/golden/tiletree_handworked.py:
the derivations are in that file).  This is what pins the restated tile tree: the fixtures are literals, not the
output of any implementation."""
import numpy as np
import pytest

from oracle import oracle
from tiletree_cases import CASES, BoxesOnly, expected_tiles, spec_of


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_tile_tree_matches_hand_derivation(name):
    case = CASES[name]
    jobs = oracle.build_raster_jobs(BoxesOnly(case), spec_of(case))
    tiles, shapes = oracle.tiles_in_tree_order(jobs)
    want, numbers = expected_tiles(case)
    assert len(tiles) == len(want)
    for field in ("left", "top", "right", "bottom", "h_depth", "v_depth", "shape_start", "shape_count"):
        assert np.array_equal(tiles[field], want[field]), (field, np.flatnonzero(tiles[field] != want[field])[:8])
    assert np.array_equal((shapes["tag"] & 0xFFFFFFFF).astype(np.int64), numbers)
    # the references keep the entry's own words (Job.hs:121-122 referenceShape)
    strands = np.asarray([s[4] for s in case["shapes"]], np.int64)
    assert np.array_equal(shapes["num_strands"].astype(np.int64), strands[numbers] if len(numbers) else numbers)
    assert np.array_equal(shapes["geo_start"].astype(np.int64), 8 * numbers)


def test_job_packing_swapped_arguments():
    case = CASES["T6_job_packing_swapped_arguments"]
    jobs = oracle.build_raster_jobs(BoxesOnly(case), spec_of(case))
    want, _ = expected_tiles(case)
    j = case["jobs"]
    assert len(jobs) == j["count"]
    for job, first in zip(jobs, j["first_leaf_of_job"]):          # last created first (CallKernels.hs:255)
        assert len(job.tiles) == j["tiles_per_job"]
        assert job.columns == j["columns_per_job"]
        assert np.array_equal(job.tiles["column_allocation"], j["column_step"] * np.arange(j["tiles_per_job"]))
        for field in ("left", "top", "right", "bottom"):
            assert np.array_equal(job.tiles[field], want[field][first:first + j["tiles_per_job"]])
    assert [len(job.shapes) for job in jobs] == [0, 0, 0, 4]
    assert np.array_equal(jobs[3].tiles["shape_start"][:5], [0, 1, 2, 3, 4])
