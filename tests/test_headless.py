"""Headless output target: PPM files written from the rasterizer's BGRA words (gudni_b200/headless.py)."""
import numpy as np
import pytest

from gudni_b200 import headless, scenes


def test_ppm_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    img = (rng.integers(0, 1 << 24, (37, 53), dtype=np.uint32)) | np.uint32(0xFF000000)
    path = headless.handle_output_ppm(str(tmp_path / "frame"), 7, img)
    assert path.endswith("frame-0007.ppm")
    with open(path, "rb") as f:
        assert f.read(11) == b"P6\n53 37\n25"
    assert np.array_equal(headless.read_ppm(path), img)
    # byte order: the word is B | G<<8 | R<<16 | A<<24 (Kernels.cl:842-844); PPM wants R, G, B
    one = np.array([[0xFF112233]], np.uint32)
    assert headless.bgra_to_rgb_bytes(one).tolist() == [[[0x11, 0x22, 0x33]]]


@pytest.mark.gpu
def test_frame_to_ppm_equals_oracle(rasterizer, tmp_path):
    from oracle import oracle
    scene = scenes.translucent_stack()
    img, _ = rasterizer.raster_scene(0, scene)
    path = headless.handle_output_ppm(str(tmp_path / "stack"), 0, img)
    assert np.array_equal(headless.read_ppm(path), oracle.render(scene, taps=False).image)
