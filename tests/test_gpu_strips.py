"""Strip partition through the C ABI on one GPU: rendering the canvas strip by strip
(gudni_b200_frame_strip) must reproduce the whole-frame image bit for bit, whether the strips
land in the context's own buffer or in a caller-owned canvas (gudni_b200_frame_target)."""
import numpy as np
import pytest

from gudni_b200 import scenes
from gudni_b200.strips import partition_rows

pytestmark = pytest.mark.gpu


def test_strips_reassemble_to_full_frame(rasterizer):
    scene = scenes.fuzzy_circles(5000, 1100, 1300, 5, 60, 0x57121)
    full, _ = rasterizer.raster_scene(0, scene)
    for n in (2, 3, 5):
        rows = partition_rows(scene, n, rasterizer.spec.max_tile_size)
        assert rows[0][0] == 0 and rows[-1][1] == scene.height
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        out = np.zeros_like(full)
        for (y0, y1) in rows:
            strip, stats = rasterizer.raster_scene(1, scene, rows=(y0, y1))
            assert strip.shape == (y1 - y0, scene.width)
            out[y0:y1] = strip
        assert np.array_equal(out, full), f"{n} strips differ from the whole frame"


def test_strip_into_caller_owned_canvas(rasterizer):
    import torch
    scene = scenes.fuzzy_circles(3000, 900, 1000, 5, 60, 0x57122)
    full, _ = rasterizer.raster_scene(0, scene)
    canvas = torch.zeros((scene.height, scene.width), dtype=torch.int32, device="cuda")
    try:
        rasterizer.frame_target(canvas.data_ptr(), 0)
        for rows in partition_rows(scene, 4, rasterizer.spec.max_tile_size):
            rasterizer.frame_begin(scene, 0)
            rasterizer.frame_strip(*rows)
            rasterizer.raster_entries(scene.subset_rows(*rows))
            rasterizer.frame_end(want_image=False)
    finally:
        rasterizer.frame_target(None)
    torch.cuda.synchronize()
    assert np.array_equal(canvas.cpu().numpy().view(np.uint32), full)


def test_bad_strip_is_rejected(rasterizer):
    from gudni_b200.raster import GudniError
    scene = scenes.tiny_square(size=600)
    rasterizer.frame_begin(scene, 0)
    with pytest.raises(GudniError):
        rasterizer.frame_strip(100, 300)   # not whole root-tile rows
    rasterizer.frame_strip(256, 512)
    rasterizer.raster_entries(scene.subset_rows(256, 512))
    rasterizer.frame_end(want_image=False)
