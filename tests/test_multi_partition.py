"""The strip partition of the multi-GPU path (gudni_b200/csrc/multi.cu), host arithmetic only: the C ABI's
gudni_b200_partition_rows / gudni_b200_rebalance_rows against the Python statement of the same rules
(gudni_b200/strips.py) and against the properties a partition must have."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.multi import partition_rows, rebalance_rows
from gudni_b200 import strips


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
def test_partition_matches_the_python_statement(n):
    scene = scenes.fuzzy_circles(4000, 1500, 2300, 5, 80, 0x9A27 + n)
    rows = partition_rows(scene.entries, scene.width, scene.height, 256, n)
    assert rows == strips.partition_rows(scene, n, 256)
    assert rows[0][0] == 0 and rows[-1][1] == scene.height
    assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert all(r[0] % 256 == 0 for r in rows)


def test_more_devices_than_tile_rows_leaves_devices_idle():
    scene = scenes.fuzzy_circles(50, 300, 500, 5, 40, 3)          # two tile rows
    rows = partition_rows(scene.entries, scene.width, scene.height, 256, 4)
    assert rows[:2] == [(0, 256), (256, 500)] and rows[2:] == [(500, 500), (500, 500)]


def test_empty_scene_is_cut_by_pixels():
    rows = partition_rows(np.zeros(0, scenes.fuzzy_circles(0, 8, 8, 1, 2, 1).entries.dtype), 1024, 2048, 256, 4)
    assert rows == [(0, 512), (512, 1024), (1024, 1536), (1536, 2048)]


@pytest.mark.parametrize("seed", range(6))
def test_rebalance_matches_the_python_statement_and_never_loses(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 9))
    height = int(rng.integers(n, 64)) * 256 - int(rng.integers(0, 200))
    n_rows = (height + 255) // 256
    cuts = sorted(rng.choice(np.arange(1, n_rows), size=n - 1, replace=False).tolist())
    bounds = [0] + cuts + [n_rows]
    rows = [(bounds[k] * 256, min(bounds[k + 1] * 256, height)) for k in range(n)]
    ms = (rng.random(n) * 5 + 0.2).tolist()
    new = rebalance_rows(rows, ms, height, 256)
    assert new == strips.rebalance_rows(rows, ms, height, 256)
    assert new[0][0] == 0 and new[-1][1] == height and all(a[1] == b[0] for a, b in zip(new, new[1:]))
    # under the model the rebalancer uses (cost uniform inside last frame's strips) the new cut finishes no later
    cost = np.zeros(n_rows)
    for (y0, y1), t in zip(rows, ms):
        a, b = y0 // 256, (y1 + 255) // 256
        cost[a:b] = t / (b - a)
    finish = lambda rr: max(cost[y0 // 256:(y1 + 255) // 256].sum() for y0, y1 in rr)
    assert finish(new) <= finish(rows) + 1e-9
