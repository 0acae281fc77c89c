"""CUDA path against the REFERENCE'S OWN KERNELS, without the restated oracle in between.

1. The committed golden vectors (tests/golden/reference_hashes.json, produced by Kernels.cl compiled for
   the host — tests/golden/make_golden.py): image, per-thread threshold counts and shape-bit counts of
   the CUDA path hash to the same values.
2. When the compiled reference travelled to this box (oracle/_ref/libgudni_ref.so is a built artefact,
   git-ignored but shipped like the other .so files), live comparisons on seeded scenes.
The raster jobs come from the restated tile tree (the reference's is Haskell); level 2 of the ABI (GPU
binning) is checked against the same vectors through the image.
"""
import json
import os

import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.formats import CANONICAL_SPEC
from gudni_b200.raster import setup_rasterizer
from oracle import oracle

pytestmark = pytest.mark.gpu

from golden.make_golden import SCENES, digest  # noqa: E402

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_hashes.json")))


class _Result:
    def __init__(self, image, counts, bits, total):
        self.image, self.n_thresholds, self.shape_bits, self.total_thresholds = image, [counts], [bits], total


def cuda_level1(r, scene, spec):
    jobs = oracle.build_raster_jobs(scene, spec)
    r.debug_enable(True)
    img, stats = r.queue_raster_jobs(0, scene, jobs)
    counts, bits = r.debug_thread_counts()
    r.debug_enable(False)
    assert stats.n_overflow_threads == 0
    return _Result(img, counts, bits, stats.n_thresholds), jobs


@pytest.fixture(scope="module")
def rasterizers():
    made = {}

    def get(spec):
        if spec not in made:
            made[spec] = setup_rasterizer(0) if spec == CANONICAL_SPEC else setup_rasterizer(0, spec)
        return made[spec]
    yield get
    for r in made.values():
        r.close()


@pytest.mark.parametrize("name", sorted(SCENES))
def test_golden_vectors_level1(rasterizers, name):
    make, spec = SCENES[name]
    spec = spec or CANONICAL_SPEC
    got, _ = cuda_level1(rasterizers(spec), make(), spec)
    assert digest(got) == GOLDEN[name], name


@pytest.mark.parametrize("name", sorted(SCENES))
def test_golden_vectors_level2_image(rasterizers, name):
    import hashlib
    make, spec = SCENES[name]
    spec = spec or CANONICAL_SPEC
    img, stats = rasterizers(spec).raster_scene(0, make())
    assert hashlib.sha256(img.astype("<u4").tobytes()).hexdigest() == GOLDEN[name]["sha256"], name
    assert stats.n_thresholds == GOLDEN[name]["thresholds"], name


needs_ref = pytest.mark.skipif(oracle.reference_lib() is None, reason="oracle/_ref/libgudni_ref.so did not travel")


def live(r, scene, spec=CANONICAL_SPEC):
    got, jobs = cuda_level1(r, scene, spec)
    ref = oracle.raster_jobs(scene, jobs, spec, reference=True)
    assert ref.overflow_threads == 0
    assert got.total_thresholds == ref.total_thresholds
    assert np.array_equal(got.n_thresholds[0], np.concatenate(ref.n_thresholds))
    assert np.array_equal(got.shape_bits[0], np.concatenate(ref.shape_bits))
    bad = np.argwhere(got.image != ref.image)
    assert len(bad) == 0, f"{len(bad)} pixels differ from the reference kernels, first at (y,x)={bad[:5].tolist()}"


@needs_ref
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_live_fuzzy_circles(rasterizers, seed):
    rng = np.random.default_rng(seed)
    w, h = int(rng.integers(200, 900)), int(rng.integers(150, 700))
    live(rasterizers(CANONICAL_SPEC), scenes.fuzzy_circles(int(rng.integers(200, 1500)), w, h, 4, 45, seed))


@needs_ref
def test_live_random_rectangles(rasterizers):
    live(rasterizers(CANONICAL_SPEC), scenes.random_rectangles(300, 640, 480, 5))


@needs_ref
def test_live_pictures(rasterizers):
    live(rasterizers(CANONICAL_SPEC), scenes.picture_scene())


@needs_ref
def test_live_s4b_full_size(rasterizers):
    """BASELINE.json's '100k curves at 3840x2160' (S4b) against the reference kernels, every pixel."""
    live(rasterizers(CANONICAL_SPEC), scenes.s4b())


OPENCL = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "opencl_b200_hashes.json")))["strict"]


@pytest.mark.parametrize("name", ["s2", "s4b", "s4"])
def test_full_size_scenes_against_the_real_opencl_run(rasterizers, name):
    """The CUDA path against what the reference's kernels produced under NVIDIA's OpenCL runtime on a
    B200 (tests/golden/opencl_b200_hashes.json, IEEE build; s2's entry comes from the reference-options
    build and is not used): same image hash, same threshold total, GPU-binned (level 2)."""
    import hashlib
    if name not in OPENCL:
        pytest.skip("not in the IEEE OpenCL run")
    img, stats = rasterizers(CANONICAL_SPEC).raster_scene(0, getattr(scenes, name)())
    assert stats.n_thresholds == OPENCL[name]["thresholds"]
    assert hashlib.sha256(img.astype("<u4").tobytes()).hexdigest() == OPENCL[name]["sha256"]


def test_golden_files_agree():
    """The vectors from the compiled-for-host reference and from the OpenCL run on the B200 are the same."""
    for name in SCENES:
        assert GOLDEN[name] == OPENCL[name], name
