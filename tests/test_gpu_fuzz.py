"""Seeded differential runs: scene families x awkward canvas sizes (not multiples of the tile size, very
wide, very tall, tiny), through level 2 of the ABI against the oracle, bit-exact.  These reach the paths
the fixed scenes do not pin individually: runs of untouched pixels next to busy columns, queues that
leave the shared-memory window, colour-cache evictions, partially covered edge tiles."""
import numpy as np
import pytest

from gudni_b200 import scenes

from parity import level2_parity

pytestmark = pytest.mark.gpu

CANVASES = [(300, 200), (1000, 37), (37, 1000), (257, 513), (64, 64), (1023, 511)]


@pytest.mark.parametrize("case", range(6))
def test_sparse_and_dense_circles(rasterizer, case):
    w, h = CANVASES[case]
    rng = np.random.default_rng(0xF00D + case)
    n = int(rng.integers(5, 400))
    rmin = float(rng.uniform(1, 8))
    rmax = rmin + float(rng.uniform(2, 120))
    level2_parity(rasterizer, scenes.fuzzy_circles(n, w, h, rmin, rmax, 0xABC0 + case))


@pytest.mark.parametrize("case", range(6))
def test_rotated_rectangles_and_cutouts(rasterizer, case):
    w, h = CANVASES[case]
    rng = np.random.default_rng(0xBEEF + case)
    n = int(rng.integers(3, 250))
    level2_parity(rasterizer, scenes.random_rectangles(n, w, h, 0x5151 + case, max_size=float(rng.uniform(6, 200))))


def test_opaque_layers_stop_the_compositing_chain(rasterizer):
    # every shape opaque: determineColor stops at the first layer (alpha == 1), stacks still differ
    level2_parity(rasterizer, scenes.random_rectangles(200, 500, 300, 77, alpha=(1.0, 1.0)))


@pytest.mark.parametrize("case", range(10))
def test_mixed_bag(rasterizer, case):
    """Free-form curves with random control points (knobs, self-intersections), slivers, holes, two-outline
    shapes, pictures at several scales — the scenes tests/test_reference_pin.py holds the oracle to the
    reference's own kernels on."""
    rng = np.random.default_rng(5000 + case)
    w, h = int(rng.integers(20, 700)), int(rng.integers(20, 500))
    level2_parity(rasterizer, scenes.mixed_bag(int(rng.integers(1, 300)), w, h, 7000 + case))
